"""Generation-time rewrite of the canonical message loop of a FLAME GPU agent script.

Every agent function of the reference's examples iterates its input messages with the same shape
(``FLAMEGPU/templates/functions.xslt`` emits it, and all 35 loops under ``examples/*/src/model/functions.c``
follow it)::

    xmachine_message_M* msg = get_first_M_message(messages, ...);
    while (msg) {
        BODY
        msg = get_next_M_message(msg, messages, ...);
    }

``get_next`` hides a two-level iteration (rows of the cell neighbourhood x messages of a row) behind a
flat cursor, so the compiler sees one loop whose back edge carries a range switch, a null test and three
phi copies of the pointer: 25-30 SASS instructions per message on sm_100a, with the range switch executed
by a few lanes at a time.  The rewrite makes the two levels explicit::

    while (msg) {
        M* const end = fgpu_range_end_M(messages); M* cur = msg;
        do { M copy = fgpu::load_message(cur); M* msg = &copy;     // one 16-byte load per quad
             { BODY }
        } while (++cur != end);
        msg = fgpu_next_range_M(messages);                          // whole warp switches rows together
    }

Same messages, same order, same arithmetic per agent: results are bit-identical (tests/test_gpu_brownian.py
runs both forms).  The inner loop is a plain counted loop the compiler unrolls and batches loads for
(measured on C2: count_neighbours 219 -> 171 us).

The default emission ("pair") goes one step further, because ptxas turns neither `unroll 2` nor a pointer
compare into tight code (two copies of the 64-bit cursor, a 64-bit compare, a spill): the body text is
instantiated three times -- once for the odd message of a range, twice in a counted loop that issues both
16-byte loads up front::

    while (msg) {
        int n = fgpu_range_count_M(messages); const M* p = msg;
        if (n & 1) { M c = fgpu::load_message(p); M* msg = &c; { BODY } ++p; }
        for (n >>= 1; n > 0; --n) { M a = load(p), b = load(p + 1);
            { M* msg = &a; { BODY } }  { M* msg = &b; { BODY } }  p += 2; }
        msg = fgpu_next_range_M(messages);
    }

The copy of BODY that stays where the author wrote it keeps its line numbers; the other two are the same text
with comments removed, folded onto the existing lines.

A loop is rewritten only when it matches exactly; anything else -- ``break``/``continue``/``goto`` in the body,
another assignment to the cursor, a different statement order -- is left alone and runs through the generic
``get_next`` (still correct, just slower).  Line numbers are preserved (all inserted text stays on existing
lines) so compiler diagnostics and ``-lineinfo`` still point at the author's file.
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import List, Optional, Tuple

_IDENT = r"[A-Za-z_]\w*"


def mask_comments_and_strings(src: str) -> str:
    """Same-length copy of `src` with comments, string and char literals blanked (newlines kept)."""
    out = list(src)
    i, n = 0, len(src)
    while i < n:
        c = src[i]
        if c == "/" and i + 1 < n and src[i + 1] == "/":
            j = src.find("\n", i)
            j = n if j < 0 else j
            for k in range(i, j):
                out[k] = " "
            i = j
        elif c == "/" and i + 1 < n and src[i + 1] == "*":
            j = src.find("*/", i + 2)
            j = n if j < 0 else j + 2
            for k in range(i, j):
                if out[k] != "\n":
                    out[k] = " "
            i = j
        elif c == '"' or c == "'":
            j = i + 1
            while j < n and src[j] != c:
                j += 2 if src[j] == "\\" else 1
            j = min(j + 1, n)
            for k in range(i + 1, j - 1):
                if out[k] != "\n":
                    out[k] = " "
            i = j
        else:
            i += 1
    return "".join(out)


def strip_comments(src: str) -> str:
    """Same-length copy of `src` with comments blanked; string and char literals are kept."""
    out = list(src)
    i, n = 0, len(src)
    while i < n:
        c = src[i]
        if c == "/" and i + 1 < n and src[i + 1] == "/":
            j = src.find("\n", i)
            j = n if j < 0 else j
            for k in range(i, j):
                out[k] = " "
            i = j
        elif c == "/" and i + 1 < n and src[i + 1] == "*":
            j = src.find("*/", i + 2)
            j = n if j < 0 else j + 2
            for k in range(i, j):
                if out[k] != "\n":
                    out[k] = " "
            i = j
        elif c == '"' or c == "'":
            j = i + 1
            while j < n and src[j] != c:
                j += 2 if src[j] == "\\" else 1
            i = min(j + 1, n)
        else:
            i += 1
    return "".join(out)


def _match_forward(text: str, pos: int, open_c: str, close_c: str) -> int:
    """Index of the bracket closing the one at `pos`; -1 if unbalanced."""
    depth = 0
    for i in range(pos, len(text)):
        if text[i] == open_c:
            depth += 1
        elif text[i] == close_c:
            depth -= 1
            if depth == 0:
                return i
    return -1


def _match_backward(text: str, pos: int, open_c: str, close_c: str) -> int:
    """Index of the bracket opening the one that closes at `pos`; -1 if unbalanced."""
    depth = 0
    for i in range(pos, -1, -1):
        if text[i] == close_c:
            depth += 1
        elif text[i] == open_c:
            depth -= 1
            if depth == 0:
                return i
    return -1


def _split_args(text: str) -> List[str]:
    args, depth, cur = [], 0, []
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            args.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    args.append("".join(cur).strip())
    return args


@dataclass
class LoopReport:
    message: str
    line: int
    rewritten: bool
    reason: str = ""


def rewrite_message_loops(src: str, message_names: List[str], unroll: Optional[int] = None,
                          style: str = "pair") -> Tuple[str, List[LoopReport]]:
    """Rewrite every canonical message loop of `src`; returns the new text and one report per get_next call.
    `style` "pair": three instances of the body (odd message + a counted loop over pairs); "single": one
    instance in a do-while over the range, with `unroll` as an unroll pragma on it (tuning knob; default: the
    compiler decides).  A body that cannot be duplicated safely (preprocessor lines, labels, static locals) is
    emitted in the single style, and so is a loop whose body holds another message walk."""
    pragma = ('_Pragma("unroll %d") ' % unroll) if unroll else ""
    masked = mask_comments_and_strings(src)
    nocomment = strip_comments(src)
    names = "|".join(re.escape(m) for m in sorted(message_names, key=len, reverse=True))
    if not names:
        return src, []
    call_re = re.compile(r"\bget_next_(%s)_message\s*\(" % names)
    edits: List[Tuple[int, int, str]] = []   # (start, end, replacement) on the original text
    reports: List[LoopReport] = []
    serial = 0
    for m in call_re.finditer(masked):
        msg = m.group(1)
        line = masked.count("\n", 0, m.start()) + 1

        def skip(reason: str) -> None:
            reports.append(LoopReport(msg, line, False, reason))

        close = _match_forward(masked, m.end() - 1, "(", ")")
        if close < 0:
            skip("unbalanced call")
            continue
        args = _split_args(masked[m.end():close])
        if len(args) < 2 or not re.fullmatch(_IDENT, args[0]) or not re.fullmatch(_IDENT, args[1]):
            skip("arguments are not plain identifiers")
            continue
        var, lst = args[0], args[1]
        # statement:  VAR = get_next_M_message(...) ;
        stmt_start = max(masked.rfind(";", 0, m.start()), masked.rfind("{", 0, m.start()), masked.rfind("}", 0, m.start())) + 1
        if not re.fullmatch(r"\s*%s\s*=\s*" % re.escape(var), masked[stmt_start:m.start()]):
            skip("not of the form `cursor = get_next(cursor, ...)`")
            continue
        tail = re.match(r"\s*;\s*", masked[close + 1:])
        if not tail:
            skip("call is not a full statement")
            continue
        stmt_end = close + 1 + tail.end()           # first char after `;` and blanks
        if stmt_end >= len(masked) or masked[stmt_end] != "}":
            skip("get_next is not the last statement of the loop body")
            continue
        body_close = stmt_end
        body_open = _match_backward(masked, body_close, "{", "}")
        if body_open < 0:
            skip("unbalanced braces")
            continue
        # loop head:  while ( VAR [!= NULL|0|nullptr] )
        head = re.search(r"\bwhile\s*\(\s*(?:%s(?:\s*!=\s*(?:NULL|0|nullptr))?|(?:NULL|0|nullptr)\s*!=\s*%s)\s*\)\s*$"
                         % (re.escape(var), re.escape(var)), masked[:body_open])
        if not head:
            skip("loop head is not `while (cursor)`")
            continue
        body = masked[body_open + 1:stmt_start]
        if re.search(r"\b(break|continue|goto)\b", body):
            skip("body contains break/continue/goto")
            continue
        v = re.escape(var)
        if re.search(r"(?<![\w.>])%s\s*(=(?!=)|\+\+|--|\+=|-=)|(\+\+|--)\s*%s\b|&\s*%s\b" % (v, v, v), body):
            skip("body modifies or takes the address of the cursor")
            continue
        if re.search(r"\bget_(first|next)_%s_message\b" % re.escape(msg), body):
            skip("nested walk over the same message list")
            continue
        serial += 1
        S = f"xmachine_message_{msg}"
        k = serial
        copy = " ".join(nocomment[body_open + 1:stmt_start].split())
        dup_ok = (style == "pair" and not re.search(r"(^|\n)\s*#", body) and not re.search(r"\bstatic\b", body)
                  and not re.search(r"\bget_next_\w+_message\b", body)      # a nested walk: its own rewrite edits this text

                  and not re.search(r"(?:^|[;{}])\s*(?!case\b|default\b)[A-Za-z_]\w*\s*:(?!:)", body))
        if dup_ok:
            prologue = (f" int fgpu_n{k} = fgpu_range_count_{msg}({lst}); const {S}* fgpu_p{k} = {var}; "
                        f"if (fgpu_n{k} & 1) {{ {S} fgpu_m{k} = fgpu::load_message(fgpu_p{k}); {S}* {var} = &fgpu_m{k}; {{ {copy} }} ++fgpu_p{k}; }} "
                        f'_Pragma("unroll 1") for (fgpu_n{k} >>= 1; fgpu_n{k} > 0; --fgpu_n{k}) {{ '
                        f"{S} fgpu_a{k} = fgpu::load_message(fgpu_p{k}); {S} fgpu_b{k} = fgpu::load_message(fgpu_p{k} + 1); "
                        f"{{ {S}* {var} = &fgpu_a{k}; {{")
            epilogue = (f"}} }} {{ {S}* {var} = &fgpu_b{k}; {{ {copy} }} }} fgpu_p{k} += 2; }} "
                        f"{var} = fgpu_next_range_{msg}({lst});")
        else:
            prologue = (f" {S}* const fgpu_end{k} = fgpu_range_end_{msg}({lst}); {S}* fgpu_cur{k} = {var}; {pragma}do {{ "
                        f"{S} fgpu_msg{k} = fgpu::load_message(fgpu_cur{k}); {S}* {var} = &fgpu_msg{k}; {{")
            epilogue = f"}} }} while (++fgpu_cur{k} != fgpu_end{k}); {var} = fgpu_next_range_{msg}({lst});"
        edits.append((body_open + 1, body_open + 1, prologue))
        # replace exactly the statement (from the cursor name to its `;`), keeping any line breaks inside
        # it, so that neither line numbers nor the author's comments move
        real_start = stmt_start + (len(masked[stmt_start:m.start()]) - len(masked[stmt_start:m.start()].lstrip()))
        real_end = close + 1 + masked[close + 1:].index(";") + 1
        edits.append((real_start, real_end, epilogue + "\n" * src[real_start:real_end].count("\n")))
        reports.append(LoopReport(msg, line, True))
    out = src
    for s, e, rep in sorted(edits, key=lambda t: (t[0], t[1]), reverse=True):
        out = out[:s] + rep + out[e:]
    return out, reports

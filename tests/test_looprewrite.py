"""The generation-time message-loop restructuring (ev_diffusion_b200/looprewrite.py): which loops it
accepts, which it must leave alone, and that line numbers never move.  The numerical proof that the
rewritten scripts compute the same thing is on the GPU side (every parity test runs rewritten builds;
test_gpu_brownian.py / test_gpu_reference_models.py also run `_generic` builds of the same scripts)."""
import os

import pytest

from common import REFERENCE, ROOT
from ev_diffusion_b200 import codegen, xmml
from ev_diffusion_b200.looprewrite import mask_comments_and_strings, rewrite_message_loops

CANON = """
__FLAME_GPU_FUNC__ int f(xmachine_memory_A* agent, xmachine_message_loc_list* loc_messages, xmachine_message_loc_PBM* pm){
    xmachine_message_loc* msg = get_first_loc_message(loc_messages, pm, agent->x, agent->y, agent->z);
    while (msg)   /* a brace in a comment } must not confuse the matcher */
    {
        if (msg->id != agent->id) { agent->n++; }   // another one {
        msg = get_next_loc_message(msg, loc_messages, pm);
    }
    return 0;
}
"""


@pytest.mark.parametrize("style", ["pair", "single"])
def test_canonical_loop_is_rewritten_and_lines_stay(style):
    out, rep = rewrite_message_loops(CANON, ["loc"], style=style)
    assert [r.rewritten for r in rep] == [True]
    assert out.count("\n") == CANON.count("\n")
    assert "get_next_loc_message" not in mask_comments_and_strings(out)
    assert "fgpu_next_range_loc(loc_messages)" in out
    if style == "single":
        assert "fgpu::load_message(fgpu_cur1)" in out
    else:
        # three instances of the body: the odd message, and a pair per trip of the counted loop; the two copies carry
        # no comments (a `//` comment folded onto one line would swallow the rest of it)
        assert out.count("agent->n++;") == 3 and out.count("another one") == 1
        assert "fgpu_range_count_loc(loc_messages)" in out and "fgpu::load_message(fgpu_p1 + 1)" in out
    # every original line other than the two edited ones is untouched
    a, b = CANON.split("\n"), out.split("\n")
    changed = [i for i, (x, y) in enumerate(zip(a, b)) if x != y]
    assert len(changed) == 2


@pytest.mark.parametrize("head", ["while (msg != NULL)", "while(msg!=0)", "while (nullptr != msg)", "while( msg )"])
def test_null_test_spellings(head):
    out, rep = rewrite_message_loops(CANON.replace("while (msg)", head), ["loc"])
    assert rep[0].rewritten


@pytest.mark.parametrize("mutation,reason", [
    (("agent->n++; }", "agent->n++; break; }"), "break"),
    (("agent->n++; }", "agent->n++; continue; }"), "break"),
    (("agent->n++; }", "msg = 0; }"), "modifies"),
    (("agent->n++; }", "helper(&msg); }"), "modifies"),
    (("msg = get_next_loc_message(msg, loc_messages, pm);", "msg = get_next_loc_message(msg, loc_messages, pm); agent->n--;"), "last statement"),
    (("msg = get_next_loc_message(msg, loc_messages, pm);", "if (agent->n < 5) msg = get_next_loc_message(msg, loc_messages, pm); else msg = 0;"), "form"),
    (("while (msg)", "while (msg && agent->n < 3)"), "loop head"),
    (("while (msg)", "for (;msg;)"), "loop head"),
])
def test_non_canonical_loops_are_left_alone(mutation, reason):
    src = CANON.replace(*mutation)
    out, rep = rewrite_message_loops(src, ["loc"])
    assert out == src
    assert rep and not any(r.rewritten for r in rep)
    assert any(reason in r.reason for r in rep), [r.reason for r in rep]


def test_two_loops_and_a_nested_pair():
    two = CANON + CANON.replace("int f(", "int g(")
    out, rep = rewrite_message_loops(two, ["loc"])
    assert [r.rewritten for r in rep] == [True, True]
    assert "fgpu_p1" in out and "fgpu_p2" in out
    nested = """
    xmachine_message_a* m = get_first_a_message(a_messages);
    while (m) {
        xmachine_message_b* q = get_first_b_message(b_messages, pm, m->x, m->y, m->z);
        while (q) {
            s += q->v * m->v;
            q = get_next_b_message(q, b_messages, pm);
        }
        m = get_next_a_message(m, a_messages);
    }
"""
    out, rep = rewrite_message_loops(nested, ["a", "b"])
    assert sorted((r.message, r.rewritten) for r in rep) == [("a", True), ("b", True)]
    assert out.count("\n") == nested.count("\n")
    # the inner loop takes the pair form; the outer one, whose body holds another walk, the single form
    assert out.count("do {") == 1 and out.count("} while (++fgpu_cur") == 1 and out.count("s += q->v * m->v;") == 3


def test_bodies_that_cannot_be_duplicated_take_the_single_form():
    for bad in ("static int seen = 0; seen++;", "again: agent->n++;", "#ifdef X\n agent->n++;\n#endif\n"):
        src = CANON.replace("agent->n++; }", bad + " }")
        out, rep = rewrite_message_loops(src, ["loc"])
        assert rep[0].rewritten and "fgpu_cur1" in out and out.count("\n") == src.count("\n"), bad


def test_message_name_that_prefixes_another():
    src = CANON.replace("loc", "location_full")
    out, rep = rewrite_message_loops(src, ["location", "location_full"])
    assert [(r.message, r.rewritten) for r in rep] == [("location_full", True)]


def test_shipped_models_every_loop_is_rewritten():
    for name in ("Brownian", "BirthDeath"):
        d = os.path.join(ROOT, "ev_diffusion_b200", "models", name, "src", "model")
        model = xmml.parse_model(os.path.join(d, "XMLModelFile.xml"))
        src = open(os.path.join(d, "functions.c")).read()
        out, rep = rewrite_message_loops(src, [m.name for m in model.messages])
        assert rep and all(r.rewritten for r in rep), [(r.line, r.reason) for r in rep]
        assert out.count("\n") == src.count("\n")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "examples")), reason="reference tree not mounted")
def test_reference_examples_every_loop_is_rewritten():
    """All message loops of the reference's continuous-agent examples have the canonical shape."""
    seen = 0
    for name in ("CirclesPartitioning_float", "CirclesBruteForce_float", "Boids_Partitioning", "Boids_BruteForce",
                 "SmoothedParticleHydrodynamics", "Keratinocyte", "StableMarriage"):
        d = os.path.join(REFERENCE, "examples", name, "src", "model")
        if not os.path.exists(os.path.join(d, "functions.c")):
            continue
        model = xmml.parse_model(os.path.join(d, "XMLModelFile.xml"))
        src = open(os.path.join(d, "functions.c"), errors="replace").read()
        out, rep = rewrite_message_loops(src, [m.name for m in model.messages])
        assert all(r.rewritten for r in rep), (name, [(r.line, r.reason) for r in rep if not r.rewritten])
        assert out.count("\n") == src.count("\n")
        seen += len(rep)
    assert seen >= 10


def test_emit_cuda_writes_rewritten_copy_and_log(tmp_path):
    d = os.path.join(ROOT, "ev_diffusion_b200", "models", "Brownian", "src", "model")
    model = xmml.parse_model(os.path.join(d, "XMLModelFile.xml"))
    ff = [os.path.join(d, "functions.c")]
    files = codegen.emit_cuda(model, ff, str(tmp_path / "on"))
    glue = open(files["model_glue.cu"]).read()
    assert "rewritten_0_functions.c" in glue and ff[0] not in glue
    assert "rewritten" in open(files["loops.txt"]).read()
    assert open(files["rewritten_0_functions.c"]).read().startswith(f'#line 1 "{ff[0]}"')
    files = codegen.emit_cuda(model, ff, str(tmp_path / "off"), rewrite_loops=False)
    assert f'#include "{ff[0]}"' in open(files["model_glue.cu"]).read()
    assert "loops.txt" not in files

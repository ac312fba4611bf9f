"""oracle/cuda_on_host/ (the host meaning of the CUDA runtime that runs the reference's generated sources on the CPU,
oracle/build_ref.py) on a small CUDA program with known answers: barriers through shared memory, atomics, symbols,
texture references, cub scan / radix sort and thrust sort_by_key; plus the launch rewriter and the barrier analysis."""
import os
import subprocess

from common import ROOT
from oracle import build_ref

SRC = os.path.join(ROOT, "tests", "cuda_on_host", "probe.cu")
HOST = os.path.join(ROOT, "oracle", "cuda_on_host")


def test_launch_rewrite_and_barrier_analysis():
    src = open(SRC).read()
    reach = build_ref.functions_reaching_barrier([src])
    assert reach == {"reverse_blocks", "block_sums"}
    out = build_ref.rewrite_launches(src, reach)
    assert "<<<" not in out
    assert "cuda_on_host::Launch(grid, block, block * sizeof(int)).run(true, [&]{ reverse_blocks(d_in, d_out); });" in out
    assert "cuda_on_host::Launch(g, b, 0, 0).run(false, [&]{ count_and_fetch(d_res, d_ring); });" in out
    # a function that only CALLS a barrier-reaching function needs fibres too; #else branches are honoured
    text = ("__device__ void leaf(){ __syncthreads(); }\n__device__ void mid(int x){ leaf(); }\n"
            "__global__ void k1(){ mid(1); }\n__global__ void k2(){ }\n"
            "#ifdef FAST\n__global__ void k3(){ if (1) {\n#else\n__global__ void k3(){ if (1) { __syncthreads();\n#endif\n } }\n")
    assert build_ref.functions_reaching_barrier([text]) == {"leaf", "mid", "k1", "k3"}
    assert build_ref.functions_reaching_barrier([text], {"FAST"}) == {"leaf", "mid", "k1"}


def test_probe_program_gives_the_cuda_answers(tmp_path):
    src = open(SRC).read()
    host = tmp_path / "probe_host.cpp"
    host.write_text("CUDA_ON_HOST_SHARED_MEMORY_DEFINITION\n" + build_ref.rewrite_launches(src, build_ref.functions_reaching_barrier([src])))
    exe = tmp_path / "probe"
    subprocess.run(["g++", "-std=c++17", "-O1", "-x", "c++", "-include", os.path.join(HOST, "cuda_on_host.h"), "-I", HOST,
                    str(host), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE, text=True).stdout.split("\n")
    # block 0 holds 1,4,...,190 reversed; block 1 likewise (193..382)
    assert out[0] == "reverse 190 1 382 193"
    sums = [sum((3 * i + 1 if i < 200 else 0) for i in range(b * 64, b * 64 + 64)) for b in range(4)]
    assert out[1] == "sums %d %d %d %d" % tuple(sums)
    # 200 increments; atomicInc wraps after 4 (200 mod 5 = 0); bind sets the offset to 0; 0.5*i*2 + cosf(0)
    assert out[2] == "fetch 200 0 0 2.0 200.0"
    assert out[3] == "scan 0 5 11 27"
    # keys mod 8 = 1,1,1,2,2,1 -> stable: indices 0,1,2,5 then 3,4
    assert out[4] == "sort 012534"
    assert out[5] == "thrust 31402"

#!/usr/bin/env python
"""What bounds the CTA-pair GEMM mainloop, and what does the vendor library reach on the same shapes?

1. cuBLAS (torch.matmul, bf16) on the score net's four GEMM shapes, alone and as the 96-GEMM chain of one step
   replayed from a CUDA graph (power-capped regime, no epilogues at all) -- the library ceiling for this step.
2. Our chain of the same 96 contractions (with their fused epilogues / attention) replayed the same way.
3. The CTA-pair kernel with half of its operand loads removed / with its epilogue removed (ldt_debug_set_gemm_mode;
   results are wrong, only the timing means anything).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from ldt_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
B = 256
M = B * 32
SHAPES = [("qkv", 3072, 1024, 1), ("fc_o", 1024, 1024, 3), ("fc1", 4096, 1024, 2), ("fc2", 1024, 4096, 3)]


def timed(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def timed_with_clocks(replay, seconds=2.0):
    """Replay for `seconds`; returns (ms per replay, median SM MHz, mean W) sampled through NVML while it runs."""
    import threading
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    ms1 = timed(replay, reps=20, warm=5)
    reps = max(20, int(seconds * 1e3 / ms1))
    samples, stop = [], threading.Event()

    def poll():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
            stop.wait(0.02)
    th = threading.Thread(target=poll)
    th.start()
    ms = timed(replay, reps=reps, warm=0)
    stop.set()
    th.join()
    samples = samples[len(samples) // 4:]   # drop the ramp
    clk = sorted(c for c, _ in samples)[len(samples) // 2] if samples else -1
    pw = sum(w for _, w in samples) / max(1, len(samples))
    return ms, clk, pw


def graph_of(fn):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def main():
    g = torch.Generator().manual_seed(0)
    ops_ = {}
    for name, N, K, epi in SHAPES:
        A = (torch.randn((M, K), generator=g) * 0.5).to(dev).bfloat16()
        # 24 distinct weights per shape, like the 24 blocks (weights stream from HBM, not L2)
        W = [(torch.randn((N, K), generator=g) / K ** 0.5).to(dev).bfloat16() for _ in range(24)]
        bias = torch.randn((N,), generator=g).to(dev)
        gate = torch.randn((1, N), generator=g).to(dev)
        dt = torch.float32 if epi == 3 else torch.bfloat16
        out = torch.zeros((M, N), dtype=dt, device=dev)
        outb = torch.zeros((M, N), dtype=torch.bfloat16, device=dev)
        ops_[name] = (A, W, bias, gate, out, outb, epi, N, K)

    def ours(name, i):
        A, W, bias, gate, out, outb, epi, N, K = ops_[name]
        kw = dict(resid=out, gate=gate, gate_stride=0, rows_per_gate=32) if epi == 3 else {}
        ops.gemm(A, W[i], bias, out, epi, backend=3, **kw)

    def cublas(name, i):
        A, W, bias, gate, out, outb, epi, N, K = ops_[name]
        torch.matmul(A, W[i].t(), out=outb)

    tf = lambda name, ms: 2.0 * M * ops_[name][7] * ops_[name][8] / ms / 1e9
    print("== isolated (CUDA events around 50 back-to-back launches, weights cycling over 24 copies)")
    for name, N, K, epi in SHAPES:
        for tag, f in (("cuBLAS", cublas), ("ours  ", ours)):
            k = [0]

            def one():
                f(name, k[0] % 24)
                k[0] += 1
            ms = timed(one)
            print(f"{name:5s} {tag}: {ms * 1e3:8.2f} us {tf(name, ms):8.1f} TFLOP/s", flush=True)

    tot_fl = sum(2.0 * M * n * k for _, n, k, _ in SHAPES) * 24
    for tag, f in (("cuBLAS", cublas), ("ours  ", ours)):
        def chain():
            for i in range(24):
                for name, *_ in SHAPES:
                    f(name, i)
        gr = graph_of(chain)
        ms, clk, pw = timed_with_clocks(gr.replay)
        print(f"chain of 96 GEMMs, graph replay, {tag}: {ms:7.3f} ms  {tot_fl / ms / 1e9:8.1f} TFLOP/s  [SM {clk} MHz, {pw:.0f} W]", flush=True)

    # fused MLP kernel against its two launches
    A1, W1s, b1, _, _, hid, _, _, _ = ops_["fc1"]
    _, W2s, b2, gate2, out2, _, _, _, _ = ops_["fc2"]
    sync = ops.mlp_sync_buffer(M, dev)
    k = [0]

    def two():
        ours("fc1", k[0] % 24)
        ours("fc2", k[0] % 24)
        k[0] += 1

    def fused():
        i = k[0] % 24
        ops.mlp(A1, W1s[i], b1, hid, W2s[i], b2, out2, sync, resid=out2, gate=gate2, gate_stride=0, rows_per_gate=32)
        k[0] += 1
    fl = tf("fc1", 1.0) + tf("fc2", 1.0)
    for tag, f, mode in (("fc1 ; fc2 (two launches)", two, 0), ("fused MLP kernel", fused, 0), ("fused MLP, counters ignored", fused, 8)):
        _lib.load().ldt_debug_set_gemm_mode(mode)
        ms = timed(f)
        _lib.load().ldt_debug_set_gemm_mode(0)
        sync.zero_()
        print(f"{tag:28s}: {ms * 1e3:8.2f} us {fl / ms:8.1f} TFLOP/s", flush=True)

    def chain_fused():
        for i in range(24):
            ours("qkv", i)
            ours("fc_o", i)
            ops.mlp(A1, W1s[i], b1, hid, W2s[i], b2, out2, sync, resid=out2, gate=gate2, gate_stride=0, rows_per_gate=32)
    gr = graph_of(chain_fused)
    ms, clk, pw = timed_with_clocks(gr.replay)
    print(f"chain with the fused MLP kernel, graph replay: {ms:7.3f} ms  {tot_fl / ms / 1e9:8.1f} TFLOP/s  [SM {clk} MHz, {pw:.0f} W]", flush=True)
    for name in ("fc1", "fc2", "qkv", "fc_o"):   # one shape at a time, sustained: clock and power each kernel settles at
        def one_shape():
            for i in range(24):
                ours(name, i)
        gr = graph_of(one_shape)
        ms, clk, pw = timed_with_clocks(gr.replay, 1.5)
        print(f"sustained 24 x {name:5s} ours  : {ms / 24 * 1e3:8.2f} us {tf(name, ms / 24):8.1f} TFLOP/s  [SM {clk} MHz, {pw:.0f} W]", flush=True)

        def one_shape_cublas():
            for i in range(24):
                cublas(name, i)
        gr = graph_of(one_shape_cublas)
        ms, clk, pw = timed_with_clocks(gr.replay, 1.5)
        print(f"sustained 24 x {name:5s} cuBLAS: {ms / 24 * 1e3:8.2f} us {tf(name, ms / 24):8.1f} TFLOP/s  [SM {clk} MHz, {pw:.0f} W]", flush=True)

    def mlp_only():
        for i in range(24):
            ops.mlp(A1, W1s[i], b1, hid, W2s[i], b2, out2, sync, resid=out2, gate=gate2, gate_stride=0, rows_per_gate=32)
    gr = graph_of(mlp_only)
    ms, clk, pw = timed_with_clocks(gr.replay, 1.5)
    print(f"sustained 24 x fused MLP   : {ms / 24 * 1e3:8.2f} us {fl / (ms / 24):8.1f} TFLOP/s  [SM {clk} MHz, {pw:.0f} W]", flush=True)

    def two_only():
        for i in range(24):
            ours("fc1", i)
            ours("fc2", i)
    gr = graph_of(two_only)
    ms, clk, pw = timed_with_clocks(gr.replay, 1.5)
    print(f"sustained 24 x (fc1 ; fc2) : {ms / 24 * 1e3:8.2f} us {fl / (ms / 24):8.1f} TFLOP/s  [SM {clk} MHz, {pw:.0f} W]", flush=True)
    if len(sys.argv) > 1 and sys.argv[1] == "none":
        return

    lib = _lib.load()
    print("== CTA-pair kernel with parts removed (timing only)")
    buf = torch.zeros((148 * 8,), dtype=torch.int64, device=dev)
    modes = ((0, "full"), (4, "no epilogue"), (8, "epi w/o TMEM ld"), (16, "epi w/o smem staging"), (32, "epi w/o global ld/st"),
             (64, "epi w/o GELU math"), (16 + 32, "epi = TMEM ld (+math) only"), (8 + 32, "epi = staging (+math) only"),
             (8 + 16, "epi = global only"), (8 + 16 + 32, "epi = math only"), (8 + 16 + 32 + 64, "epi = nothing but bias"),
             (1, "no A loads"), (0, "full"))
    if len(sys.argv) > 1:
        modes = tuple((int(m), f"mode {m}") for m in sys.argv[1].split(","))
    for mode, tag in modes:
        lib.ldt_debug_set_gemm_mode(mode)
        for name, N, K, epi in SHAPES:
            if name == "qkv":
                continue
            k = [0]

            def one():
                ours(name, k[0] % 24)
                k[0] += 1
            ms = timed(one)
            lib.ldt_debug_set_gemm_counters(buf.data_ptr())
            ours(name, 0)
            torch.cuda.synchronize()
            lib.ldt_debug_set_gemm_counters(None)
            c = buf.view(148, 8).cpu().double()
            lead = c[0::2]
            print(f"{tag:30s} {name:5s}: {ms * 1e3:8.2f} us {tf(name, ms):8.1f} TFLOP/s | MMA thread {lead[:, 0].mean():.0f} clk, "
                  f"wait TMA {lead[:, 1].mean():.0f}, wait acc {lead[:, 2].mean():.0f}, tiles {lead[:, 7].mean():.2f}, "
                  f"max-tile pair MMA {lead[:, 0].max():.0f}", flush=True)
    lib.ldt_debug_set_gemm_mode(0)


if __name__ == "__main__":
    main()

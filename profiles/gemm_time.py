"""Time the node-side GEMM shapes of the C4 step (1 M atoms, 250 k per element) one by one: ms, algorithmic GB/s, TFLOP/s
(3 tf32 MMAs per product).  Run on the GPU box: python profiles/gemm_time.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hermnet_b200 import ops  # noqa: E402

SHAPES = [  # (M, K, N, mode, out2, label)
    (1_000_000, 128, 512, 1, True, "x_proj[0] all sub-networks (+pre)"),
    (1_000_000, 128, 384, 0, False, "x_proj[2] one sub-network"),
    (750_000, 128, 256, 0, False, "vec_proj"),
    (250_000, 256, 128, 1, True, "xvec_proj[0] (+pre)"),
    (250_000, 128, 384, 0, False, "xvec_proj[2]"),
    (250_000, 384, 128, 2, False, "bwd xvec_proj[2]"),
    (250_000, 128, 256, 0, False, "bwd xvec_proj[0]"),
    (750_000, 256, 128, 0, False, "bwd vec_proj"),
    (1_000_000, 384, 128, 2, False, "bwd x_proj[2]"),
    (1_000_000, 512, 128, 0, False, "bwd x_proj[0]"),
]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    dev = torch.device("cuda", 0)
    tot = 0.0
    for M, K, N, mode, two, label in SHAPES:
        a = torch.randn(M, K, device=dev)
        w = torch.randn(N, K, device=dev) / K ** 0.5
        hi, lo = ops.split_tf32(w)
        bias = torch.randn(N, device=dev) if mode != 2 else None
        out = torch.empty(M, N, device=dev)
        out2 = torch.empty(M, N, device=dev) if two else None
        aux = torch.randn(M, N, device=dev) if mode == 2 else None
        for _ in range(3):
            ops.gemm_tf32x3_ex(a, hi, lo, bias, out=out, mode=mode, aux=aux, out2=out2)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(reps):
            ops.gemm_tf32x3_ex(a, hi, lo, bias, out=out, mode=mode, aux=aux, out2=out2)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        gb = 4.0 * (M * K + M * N * (1 + (1 if two else 0) + (1 if mode == 2 else 0))) / 1e9
        tf = 3 * 2.0 * M * K * N / 1e12
        tot += ms
        print(f"{label:36s} {M:>8d} x {K:3d} -> {N:3d} mode {mode}: {ms:7.3f} ms  {gb / ms * 1e3:7.0f} GB/s  {tf / ms * 1e3:6.0f} TF/s", flush=True)
    print(f"sum {tot:.3f} ms")


if __name__ == "__main__":
    main()

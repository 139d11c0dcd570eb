"""The TF32 and bf16 matmul rates cuBLAS reaches on this box (square 8192^3, CUDA events): the practical tensor-core
ceilings the 3xTF32 matrix DFT is compared with in DESIGN 4.3 (its roofline object uses MEASURED_PEAKS.json's bf16 figure)."""
import torch
n = 8192
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, dt, tf32 in (('tf32', torch.float32, True), ('bf16', torch.bfloat16, False), ('fp32 (no tensor cores)', torch.float32, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn((n, n), device='cuda', dtype=dt); b = torch.randn((n, n), device='cuda', dtype=dt)
    reps = 3 if name.startswith('fp32') else 20
    for _ in range(2): (a @ b)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps): (a @ b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f'cuBLAS {name}: {2 * n ** 3 / ms / 1e9:.0f} TFLOP/s ({ms:.3f} ms per {n}^3)')

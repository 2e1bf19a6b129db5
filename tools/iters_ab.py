"""How much of the K1 kernels is the Sinkhorn chain?  Alternates runs of 20 training steps (bench.py's timed region) with
different iteration counts / the adaptive flag, with a pause between runs so that every run starts from the same clocks
(back-to-back timing loops drift by 10-15 % with the power cap: an earlier sweep, tools/fwd_iters.py, mistook that drift
for the cost of the chain).  Prints the per-kernel event times (hvs_mhc_stream_profile hooks)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, hvs_b200
T = 1 << 20
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1234)
x = torch.randn(T, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
dy = torch.randn(T, 4, 512, generator=g, device=dev, dtype=torch.bfloat16)
gp = torch.Generator(device=dev).manual_seed(0)
phi = torch.randn(2048, 24, generator=gp, device=dev) * 0.02
bias = torch.zeros(24, device=dev); alpha = torch.full((3,), 0.01, device=dev); scale = torch.ones(2048, device=dev)
y = torch.empty_like(x); dx = torch.empty_like(x); saved = hvs_b200.ops.new_saved(x)
lib = hvs_b200.load_library()
ws = torch.empty(int(lib.hvs_mhc_stream_bwd_saved_workspace(T, 4, 512)), dtype=torch.uint8, device=dev)
def run(iters, adaptive, steps=20):
    def step():
        hvs_b200.ops.mhc_stream_fwd(x, phi, bias, alpha, scale, iters, 1e-8, 1e-8, out=y, saved=saved, adaptive=adaptive)
        hvs_b200.ops.mhc_stream_bwd_saved(x, dy, saved, phi, bias, alpha, scale, iters, 1e-8, 1e-8, out=dx, workspace=ws, adaptive=adaptive)
    for _ in range(5): step()
    torch.cuda.synchronize()
    lib.hvs_mhc_stream_profile(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): step()
    b.record(); torch.cuda.synchronize()
    buf = (ctypes.c_float * 4)()
    lib.hvs_mhc_stream_kernel_ms(buf); lib.hvs_mhc_stream_profile(0)
    return a.elapsed_time(b) / steps, buf[0], buf[1]
for rep in range(2):
    for iters, adaptive in ((20, False), (0, False), (20, True), (5, False), (20, False)):
        time.sleep(1.5)
        s, f, bw = run(iters, adaptive)
        print(f"rep {rep} iters {iters:2d} adaptive {int(adaptive)}: step {s:.3f} ms  fwd {f:.3f}  bwd {bw:.3f}", flush=True)

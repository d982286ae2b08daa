"""cuFFT plan shapes for the 1024^3 R2C transform (GPU box): the in-place 64-bit plan of libabk against torch's out-of-place
rfftn and a 2-D + 1-D decomposition."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from abacusutils_b200._lib import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = Engine.get()
ldz = 2 * (n // 2 + 1)


def timeit(f, reps=5):
    f(); f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


g = torch.rand((n, n, ldz), device='cuda', dtype=torch.float32)
print('libabk in-place R2C (64-bit plan, caller work area): %.2f ms' % timeit(lambda: eng.rfft3_inplace(g, n, n, n)))
x = torch.rand((n, n, n), device='cuda', dtype=torch.float32)
print('torch.fft.rfftn out of place, contiguous input:       %.2f ms' % timeit(lambda: torch.fft.rfftn(x)))
xp = g[:, :, :n]
print('torch.fft.rfftn out of place, padded (strided) input: %.2f ms' % timeit(lambda: torch.fft.rfftn(xp)))
print('torch 2-D rfft2 over (y,z) + 1-D fft over x:          %.2f ms' % timeit(lambda: torch.fft.fft(torch.fft.rfft2(x), dim=0)))
y = torch.fft.rfft2(x)
print('   of which the 1-D pass over x (strided):            %.2f ms' % timeit(lambda: torch.fft.fft(y, dim=0)))
print('   of which the 2-D pass:                             %.2f ms' % timeit(lambda: torch.fft.rfft2(x)))
print('libabk 3-D plan work area: %d bytes' % eng.rfft3_plan(n, n, n)[1])

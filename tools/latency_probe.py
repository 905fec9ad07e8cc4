"""kernel time vs batch size for the two cfg2 kernels (device-resident, CUDA events)"""
import sys
sys.path.insert(0, '.')
import torch
from honeybadgermpc_b200 import _native, ntl
from honeybadgermpc_b200.field import GF
from honeybadgermpc_b200.polynomial import EvalPoint
P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
pt = EvalPoint(GF(P), 16, True)
omega = ntl.pack_vec([pt.omega.value], P)[0]
ctx = _native.get_context(P, 0)
st = torch.cuda.Stream(); ctx.set_stream(st.cuda_stream)
ZS = [1, 3, 4, 9, 12, 15]
with torch.cuda.stream(st):
    for batch in [32, 1024, 8192, 16384, 32768, 65536, 98304, 131072, 262144]:
        c = torch.randint(0, 2**62, (batch, 6, 4), dtype=torch.int64, device='cuda')
        e = torch.empty((batch, 16, 4), dtype=torch.int64, device='cuda')
        r = torch.empty((batch, 6, 4), dtype=torch.int64, device='cuda')
        res = []
        for fn in (lambda: ctx.fft_batch_evaluate(omega, 16, c.data_ptr(), batch, 6, 16, e.data_ptr(), 1),
                   lambda: ctx.fft_batch_interpolate(omega, 16, ZS, c.data_ptr(), batch, r.data_ptr(), 1)):
            for _ in range(5): fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            st.synchronize(); a.record(st)
            for _ in range(200): fn()
            b.record(st); st.synchronize()
            res.append(a.elapsed_time(b) / 200 * 1000)
        print(batch, 'encode %.1f us  interp %.1f us   per-64k: %.1f %.1f' % (res[0], res[1], res[0]*65536/batch, res[1]*65536/batch))

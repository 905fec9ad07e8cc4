"""host-side cost of one C-ABI call (enqueue only) vs GPU time"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from honeybadgermpc_b200 import _native, ntl
from honeybadgermpc_b200.field import GF
from honeybadgermpc_b200.polynomial import EvalPoint
P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
pt = EvalPoint(GF(P), 16, True)
omega = ntl.pack_vec([pt.omega.value], P)[0]
ctx = _native.get_context(P, 0)
st = torch.cuda.Stream(); ctx.set_stream(st.cuda_stream)
ZS = [1, 3, 4, 9, 12, 15]
zs32 = np.ascontiguousarray(ZS, dtype=np.int32)
with torch.cuda.stream(st):
    for batch in [32, 65536]:
        c = torch.randint(0, 2**62, (batch, 6, 4), dtype=torch.int64, device='cuda')
        e = torch.empty((batch, 16, 4), dtype=torch.int64, device='cuda')
        r = torch.empty((batch, 6, 4), dtype=torch.int64, device='cuda')
        cp, ep, rp = c.data_ptr(), e.data_ptr(), r.data_ptr()
        for name, fn in (("encode", lambda: ctx.fft_batch_evaluate(omega, 16, cp, batch, 6, 16, ep, 1)),
                         ("interp", lambda: ctx.fft_batch_interpolate(omega, 16, zs32, cp, batch, rp, 1))):
            for _ in range(5): fn()
            st.synchronize()
            t0 = time.perf_counter()
            for _ in range(300): fn()
            t1 = time.perf_counter()
            st.synchronize()
            t2 = time.perf_counter()
            print(batch, name, 'host enqueue %.1f us/call, total %.1f us/call' % ((t1 - t0) / 300 * 1e6, (t2 - t0) / 300 * 1e6))

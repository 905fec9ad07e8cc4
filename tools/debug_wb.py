import random, sys
sys.path.insert(0, '.')
from oracle import hbmpc_oracle as orc
from honeybadgermpc_b200 import robust, ntl
P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
rng = random.Random(1)
for (p, n, k, nerr, nerase) in [(53, 5, 2, 0, 0), (53, 5, 2, 1, 0), (53, 7, 3, 0, 2), (53, 7, 3, 1, 1), (P, 7, 3, 1, 1), (53, 22, 8, 0, 0), (53, 22, 8, 3, 0), (P, 16, 6, 5, 0)]:
    t = k - 1
    xs = list(range(1, n + 1))
    msg = [rng.randrange(p) for _ in range(k)]
    word = [orc.poly_eval(msg, x, p) for x in xs]
    for i in rng.sample(range(n), nerr):
        word[i] = (word[i] + 1) % p
    keep = sorted(rng.sample(range(n), n - nerase))
    pts = [xs[i] for i in keep]
    row = [word[i] for i in keep]
    e_max = (n - nerase - t) // 2
    c, ln, st = robust.wb_decode_batch_limbs(ntl.pack_vec(pts, p), ntl.pack_rows([row], len(row), p), k, e_max, p)
    got = ntl.unpack_rows(c)[0][:ln[0]]
    ext = [None] * n
    for i, v in zip(keep, row):
        ext[i] = v
    try:
        want = orc.wb_decode(ext, n, k, p, lambda i: xs[i])
    except Exception as e:
        want = repr(e)
    print(p == P, n, k, nerr, nerase, 'e_max', e_max, 'status', st[0], 'len', ln[0], 'got', got[:4], 'want', want if isinstance(want, str) else want[:4], 'msg', msg[:4])

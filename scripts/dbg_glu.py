import torch, sys
sys.path.insert(0, '.')
from multimodalanalytical_b200 import ops
from multimodalanalytical_b200._lib import *
DEV='cuda'
def _rand(*shape, dtype=torch.float32, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)
M,N,K = 25472,3072,768
dt=torch.bfloat16
h = _rand(M, K, dtype=dt)
W1, Wg = _rand(N, K, dtype=dt, scale=K ** -0.5, seed=1), _rand(N, K, dtype=dt, scale=K ** -0.5, seed=2)
b1, bg = _rand(N, seed=3), _rand(N, seed=4)
a, z1, z2 = (torch.empty(M, N, device=DEV, dtype=dt) for _ in range(3))
assert ops.ffn_glu_fwd(h, W1, Wg, b1, bg, M, N, K, a, z1=z1, z2=z2)
ad, zs, ad2, z2b = (torch.zeros(M, N, device=DEV, dtype=dt) for _ in range(4))
assert ops.ffn_glu_fwd(h, W1, Wg, b1, bg, M, N, K, ad, p_drop=0.1, seed=321, site=9)
ops.gemm(h, W1, M, N, K, ops.make_epi(EPI_STORE, zs, bias=b1), max_ctas=148)
ops.gemm(h, Wg, M, N, K, ops.make_epi(EPI_GLU_MUL, ad2, out2=z2b, bias=bg, aux=zs, p_drop=0.1, seed=321, site=9), max_ctas=148)
dy = _rand(M, K, dtype=dt, seed=5)
W2 = _rand(K, N, dtype=dt, scale=N ** -0.5, seed=6)
e1, e2, f1, f2 = (torch.zeros(M, N, device=DEV, dtype=dt) for _ in range(4))
assert ops.ffn_dglu(dy, W2, M, N, K, z1, z2, e1, e2, p_drop=0.1, seed=321, site=9, drop_ld=N)
ops.gemm(dy, W2, M, N, K, ops.make_epi(EPI_DGLU, f1, out2=f2, aux=z1, aux2=z2, p_drop=0.1, seed=321, site=9, drop_ld=N), b_mn=True, max_ctas=148)
torch.cuda.synchronize()
def cmp(n, x, y):
    d = ((x == 0) != (y == 0))
    idx = d.nonzero()
    print(n, 'mismatches', int(d.sum()), 'first', idx[:8].tolist(), 'zeros', int((x==0).sum()), int((y==0).sum()))
    for r, c in idx[:6].tolist():
        print('   ', r, c, float(x[r, c]), float(y[r, c]), 'z1', float(z1[r, c]), 'z2', float(z2[r, c]), 'a', float(a[r,c]))
cmp('ad vs ad2', ad, ad2); cmp('e2 vs f2', e2, f2); cmp('e2 vs ad', e2, ad); cmp('e1 vs ad', e1, ad); cmp('f2 vs ad2', f2, ad2)

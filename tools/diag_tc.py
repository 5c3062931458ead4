import ctypes as C, sys
sys.path.insert(0, '.')
import numpy as np, torch
from oracle import ppde_port as port
from ppde_b200 import _lib
from ppde_b200.engine import PoEModel, _ptr, _stream
def dec(mk,n,nets,J2):
    k = mk.cpu().numpy().astype(np.uint64).reshape(n,nets,J2)
    return (k>>np.uint64(32)).astype(np.uint32).view(np.float32), (np.uint64(0xFFFFFFFF)-(k&np.uint64(0xFFFFFFFF))).astype(np.int64)
for L,n in [(40,7),(96,33),(104,20),(238,19)]:
    w = port.synthetic_weights(L, seed=L, lamda=1.0)
    m = PoEModel(w.wt, w.J, w.h, w.win_lo, w.cnn, w.lamda, device="cuda:0")
    rng = np.random.default_rng(L)
    aa = rng.integers(0,20,size=(n,L)).astype(np.uint8)
    pad = np.zeros((n, m.aa_stride), dtype=np.uint8); pad[:, :L] = aa
    aad = torch.from_numpy(pad).to(m.device)
    J2, nets = 2*m.C, m.n_nets
    mk1 = torch.zeros(n*nets*J2, dtype=torch.int64, device=m.device)
    mk2 = torch.full((n*nets*J2,), -1, dtype=torch.int64, device=m.device)
    _lib.check(m.lib.ppde_cnn_forward(C.byref(m.cnn), _ptr(aad), m.aa_stride, n, _ptr(mk1), _stream()), "simt")
    _lib.check(m.lib.ppde_cnn_forward_tc(C.byref(m.cnn), _ptr(aad), m.aa_stride, n, _ptr(mk2), None, None, _stream()), "tc")
    torch.cuda.synchronize()
    v1,p1 = dec(mk1,n,nets,J2); v2,p2 = dec(mk2,n,nets,J2)
    # exact fp64 reference
    x = port.aa_to_onehot(aa).double()
    ref = np.zeros_like(v1, dtype=np.float64)
    for k,net in enumerate(w.cnn):
        z = torch.relu(torch.nn.functional.conv1d(x.transpose(1,2), torch.from_numpy(net["W0"]).double(), torch.from_numpy(net["b0"]).double()).transpose(1,2))
        r2 = torch.relu(torch.nn.functional.linear(z, torch.from_numpy(net["W1"]).double(), torch.from_numpy(net["b1"]).double()))
        ref[:,k,:] = r2.max(1)[0].numpy()
    mx = np.abs(ref).max()
    print(f"L={L}: max|v|={mx:.3f}  simt abs err {np.abs(v1-ref).max():.2e}  tc abs err {np.abs(v2-ref).max():.2e}  tc mean abs err {np.abs(v2-ref).mean():.2e}  argmax diff {np.mean(p1!=p2):.4f}")
    bad = np.argwhere(np.abs(v2-ref) > 1e-4*mx)
    print("   n bad", len(bad), bad[:10].tolist())

import ctypes as C, sys, torch
sys.path.insert(0, ".")
from f16_jsb_b200 import _lib
L = _lib.load()
k, n = int(sys.argv[1]) if len(sys.argv) > 1 else 128, int(sys.argv[2]) if len(sys.argv) > 2 else 32
def run(x, dy):
    dw = torch.zeros((n, k), device="cuda"); db = torch.zeros((n,), device="cuda")
    _lib.check(L.f16_lma_linear_wgrad_tc(x.shape[0], k, n, C.c_void_p(x.data_ptr()), C.c_void_p(dy.data_ptr()), C.c_void_p(dw.data_ptr()), C.c_void_p(db.data_ptr()), None), "wg")
    torch.cuda.synchronize()
    return dw, db
for (m0, n0, k0) in [(0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 5, 9), (1, 0, 0), (3, 2, 7), (9, 4, 40), (0, 0, 33), (0, 31, 127), (15, 17, 100)]:
    x = torch.zeros((16, k), device="cuda"); dy = torch.zeros((16, n), device="cuda")
    x[m0, k0] = 1.0; dy[m0, n0] = 1.0
    dw, db = run(x, dy)
    nz = dw.nonzero().tolist()
    print("one-hot m=%d n=%d k=%d -> nonzeros %s values %s | db nz %s" % (m0, n0, k0, nz[:6], [round(float(dw[i, j]), 4) for i, j in nz[:6]], db.nonzero().flatten().tolist()))
# a dense row: x[0, :] = arange, dy[0, n0] = 1
x = torch.zeros((16, k), device="cuda"); dy = torch.zeros((16, n), device="cuda")
x[0] = torch.arange(k, device="cuda").float() + 1; dy[0, 3] = 1.0
dw, _ = run(x, dy)
print("row 3 of dw:", dw[3, :40].tolist())
print("other rows nonzero:", (dw.abs().sum(1) > 0).nonzero().flatten().tolist())

import ctypes as C, os, sys, torch
sys.path.insert(0, ".")
k, n, rows = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 655360
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
os.environ["F16_WG_DBG"] = str(dbg.data_ptr())
from f16_jsb_b200 import _lib
L = _lib.load()
x = torch.randn((rows, k), device="cuda"); dy = torch.randn((rows, n), device="cuda")
dw = torch.zeros((n, k), device="cuda"); db = torch.zeros((n,), device="cuda")
for _ in range(3):
    _lib.check(L.f16_lma_linear_wgrad_tc(rows, k, n, C.c_void_p(x.data_ptr()), C.c_void_p(dy.data_ptr()), C.c_void_p(dw.data_ptr()), C.c_void_p(db.data_ptr()), None), "wg")
torch.cuda.synchronize()
d = dbg.tolist()
for g in (0, 1):
    it = max(1, d[8 * g + 4])
    print("converter group %d: chunks %d; per chunk cycles: wait raw_full %.0f, wait empty %.0f, load+convert+store %.0f, fence+arrive %.0f" % (g, it, d[8*g]/it, d[8*g+1]/it, d[8*g+2]/it, d[8*g+3]/it))
it = max(1, d[18])
print("TMA producer: chunks %d; per chunk: wait raw_empty %.0f, total %.0f" % (it, d[16]/it, d[17]/it))
print("MMA warp: per chunk: wait acc_empty %.0f, wait full %.0f, issue %.0f, total %.0f" % (d[24]/it, d[25]/it, d[26]/it, d[27]/it))

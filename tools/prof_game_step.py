"""K1 alone on 16 Mi device-resident states (for ncu)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import corintho_ai_b200 as cb
L = cb.lib(); n = 1 << 24; dev = torch.device("cuda", 0)
a = torch.zeros((n, 2), dtype=torch.int64, device=dev); a[:, 1] = 0x0000040404040404
b = torch.empty_like(a); mf = torch.empty((n, 4), dtype=torch.int32, device=dev)
for r in range(14):
    assert L.cb200_game_step_device(n, C.c_void_p(a.data_ptr()), 1000 + r, C.c_void_p(mf.data_ptr()), C.c_void_p(b.data_ptr()), None) == 0
    a, b = b, a
torch.cuda.synchronize(); print("ok")

"""K1 at the bench size (64 Mi device-resident reachable positions): the paired kernel (default;
CB200_K1_PAIR selects its other forms) against the split and the select-only kernels -- outputs compared on the device for every
position, then every variant timed like bench.measure_game_logic (10 launches, CUDA events)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import corintho_ai_b200 as cb

L = cb.lib()
L.cb200_set_device(0)
dev = torch.device("cuda", 0)
n = 1 << 26
a = torch.zeros((n, 2), dtype=torch.int64, device=dev)
a[:, 1] = 0x0000040404040404
b = torch.empty_like(a)
mf = torch.empty((n, 4), dtype=torch.int32, device=dev)


def step(src, seed, m, dst):
    assert L.cb200_game_step_device(n, C.c_void_p(src.data_ptr()), seed, C.c_void_p(m.data_ptr()),
                                    C.c_void_p(dst.data_ptr()), None) == 0


def variant(env):
    for k in ("CB200_K1_SPLIT", "CB200_K1_SELECT_ONLY", "CB200_K1_WAVES", "CB200_K1_PAIR"):
        os.environ.pop(k, None)
    os.environ.update(env)


for r in range(12):
    step(a, 1000 + r, mf, b)
    a, b = b, a
torch.cuda.synchronize()
out = {"states": n}
# parity at full size: same outputs from all three kernels, three rounds of play
b2, mf2 = torch.empty_like(a), torch.empty_like(mf)
same = True
for r in range(3):
    variant({})
    step(a, 500 + r, mf, b)
    for env in ({"CB200_K1_SPLIT": "1"}, {"CB200_K1_SELECT_ONLY": "1"}, {"CB200_K1_WAVES": "1"}, {"CB200_K1_PAIR": "1"},
                {"CB200_K1_PAIR": "2"}, {"CB200_K1_PAIR": "3"}, {"CB200_K1_PAIR": "4"}, {"CB200_K1_PAIR": "5"},
                {"CB200_K1_PAIR": "6"}, {"CB200_K1_PAIR": "0"}):
        variant(env)
        step(a, 500 + r, mf2, b2)
        torch.cuda.synchronize()
        ok = bool(torch.equal(mf, mf2)) and bool(torch.equal(b, b2))
        same = same and ok
        if not ok:
            bad = int(((mf != mf2).any(1) | (b != b2).any(1)).sum())
            out.setdefault("mismatch", []).append({"round": r, "env": env, "positions": bad})
    a, b = b, a
out["every_form_equals_the_default"] = same
del b2, mf2
variant({})
timings = {}
forms = [("pair", {})]
for f, ws in (("0", ("8",)), ("1", ("8",)), ("5", ("8",)), ("6", ("4", "8", "16")), ("2", ("8",)), ("4", ("8",))):
    for w in ws:
        forms.append(("pair_form%s_waves%s" % (f, w), {"CB200_K1_PAIR": f, "CB200_K1_WAVES": w}))
forms += [("split", {"CB200_K1_SPLIT": "1"}), ("select_only", {"CB200_K1_SELECT_ONLY": "1"}), ("pair_again", {})]
for name, env in forms:
    variant(env)
    step(a, 7, mf, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(10):
        step(a, 77 + r, mf, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    timings[name] = {"ms_per_launch": round(ms, 4), "states_per_sec": n / (ms * 1e-3), "GBps_algorithmic_46B": 46.0 * n / (ms * 1e-3) / 1e9}
out["timings"] = timings
print(json.dumps(out))

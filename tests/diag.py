"""Step-by-step engine-vs-oracle comparison used to localise a parity failure (test infra)."""
import numpy as np


def _records(words):
    """Split an arena dump into records; returns list of (offset, header, slots[n,4])."""
    out, off = [], 0
    while off + 8 <= len(words):
        n = int(words[off + 4] & 0xFF)
        out.append((off, words[off:off + 8], words[off + 8:off + 8 + 4 * n].reshape(n, 4)))
        off += 8 + 4 * n
    return out


def _rec(words, off):
    n = int(words[off + 4] & 0xFF)
    return words[off:off + 8], words[off + 8:off + 8 + 4 * n].reshape(n, 4)


def compare_trees(eng, orc_lib, orc_tr, game, player):
    """Logical comparison of one tree of the engine with the oracle's: same statistics, priors
    and shape from the root down (record offsets may differ: the engine re-roots in place).
    Returns '' if equal else a description of the first difference."""
    import ctypes as C
    eo, ew = eng.dump_tree(game, player)
    oo = np.zeros(8, np.int64)
    ow = np.zeros(1 << 22, np.uint32)
    f = orc_lib.lib.orc_trainer_dump_tree
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    used = f(orc_tr.h, game, player, oo.ctypes.data_as(C.c_void_p), ow.ctypes.data_as(C.c_void_p), len(ow))
    ow = ow[:used]
    e_root = int(eo[7]) >> 8
    eo = eo.copy()
    eo[7] &= 0xFF
    names = ["has_root", "used", "root_visits", "root_result", "root_allv", "searches_done", "root_eval_bits", "to_play"]
    if not eo[0] and not oo[0]:
        return ""
    for i, nm in enumerate(names):
        if nm != "used" and eo[i] != oo[i]:
            return f"game {game} player {player}: ctl {nm}: engine {eo[i]} oracle {oo[i]} (all: {eo} vs {oo})"
    stack = [(e_root, 0, "root")]
    keep = np.uint32(0xFFFFFFFF & ~(1 << 28))  # engine-only child_has_children bit
    while stack:
        eoff, ooff, path = stack.pop()
        h1, s1 = _rec(ew, eoff)
        h2, s2 = _rec(ow, ooff)
        if h1[4] != h2[4] or h1[5] != h2[5]:
            return f"game {game} player {player}: node {path}: header {h1} vs {h2}"
        a, b = s1.copy(), s2.copy()
        a[:, 3] &= keep
        a[:, 2] = 0
        b[:, 2] = 0
        if not (a == b).all():
            k = int(np.argwhere((a != b).any(1))[0][0])
            return (f"game {game} player {player}: node {path} depth {(h1[4] >> 8) & 255} slot {k}: engine {a[k]} "
                    f"(eval {a[k][0:1].view(np.float32)[0]}) oracle {b[k]} (eval {b[k][0:1].view(np.float32)[0]})")
        for k in range(len(s1)):
            if (s1[k][3] >> 20) & 1:
                stack.append((int(s1[k][2]), int(s2[k][2]), path + "/" + str(int(s1[k][3] & 0x7F))))
    return ""


def first_divergence(make_engine, make_oracle, orc_lib, evaluator, testing=False, max_rounds=100000):
    """Run both side by side; return a message describing the first differing round."""
    e, o = make_engine(), make_oracle()
    G, spe = e.num_games, e.spe
    ev = np.zeros(G * spe, np.float32)
    pr = np.zeros((G * spe, 96), np.float32)
    tp = 0 if testing else -1
    for rnd in range(max_rounds):
        de, do = e.do_iteration(ev, pr, tp), o.do_iteration(ev, pr, tp)
        for g in range(G):
            for p in range(2):
                msg = compare_trees(e, orc_lib, o, g, p)
                if msg:
                    return f"round {rnd}: {msg}"
        if de != do:
            return f"round {rnd}: done flag engine {de} oracle {do}"
        if de:
            return ""
        ne, no = e.num_requests(tp), o.num_requests(tp)
        if ne != no:
            return f"round {rnd}: num_requests engine {ne} oracle {no}"
        if ne == 0:
            tp = 1 - tp
            continue
        re_, ro = e.write_requests(tp), o.write_requests(tp)
        if re_.tobytes() != ro.tobytes():
            bad = int(np.argwhere((re_ != ro).any(1))[0][0])
            return f"round {rnd}: request row {bad} differs: engine {re_[bad]} oracle {ro[bad]}"
        a, b = evaluator(ro)
        ev[:no], pr[:no] = a, b
    return "no divergence within max_rounds"

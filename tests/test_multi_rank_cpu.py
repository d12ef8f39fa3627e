"""N > 1 host logic on CPU: world_size-2 gloo run of the shard partition + sample all-gather.
Each rank plays its shard's games with the CPU oracle standing in for the GPU engine (the
gather code is device-agnostic) and the gathered rows must equal the single-run order."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CFG = dict(num_games=7, seed=99, max_searches=24, searches_per_eval=8, c_puct=1.0, epsilon=0.25)


def identity_rows(oracle_lib):
    """Run all games with the oracle; return per-game lists of (state70, probs96, label)."""
    from oracle.pyoracle import synth_eval
    from util import run_trainer
    t = oracle_lib.trainer(**CFG)
    r = run_trainer(t, synth_eval)
    gs, ev, pr = r["samples"]
    return gs[::8], pr[::8], ev[::8], r


def per_game_counts(oracle_lib):
    """Samples per game = run each game count prefix (games are independent)."""
    from oracle.pyoracle import synth_eval
    from util import run_trainer
    counts = []
    prev = 0
    for g in range(1, CFG["num_games"] + 1):
        cfg = dict(CFG, num_games=g)
        t = oracle_lib.trainer(**cfg)
        # staggering depends on num_games / max_searches only through max(.,1) here (7 < 24)
        r = run_trainer(t, synth_eval)
        counts.append(r["num_samples"] - prev)
        prev = r["num_samples"]
    return counts


def worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from corintho_ai_b200.dist import all_gather_rows, shard_range
    from oracle.pyoracle import OracleLib
    O = OracleLib()
    gs, pr, ev, _ = identity_rows(O)
    counts = per_game_counts(O)
    starts = np.concatenate([[0], np.cumsum(counts)])
    first, cnt = shard_range(CFG["num_games"], world, rank)
    lo, hi = starts[first], starts[first + cnt]
    rows = np.concatenate([gs[lo:hi], pr[lo:hi], ev[lo:hi, None]], 1).astype(np.float32)
    allrows, cnts = all_gather_rows(dist, rows, torch.device("cpu"))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), allrows)
    np.save(os.path.join(out_dir, f"counts{rank}.npy"), np.array(cnts))
    dist.destroy_process_group()


def test_shard_range_partitions_every_game_once():
    from corintho_ai_b200.dist import shard_range
    for total in (1, 7, 8, 4096, 10000):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                f, c = shard_range(total, world, r)
                seen += list(range(f, f + c))
            assert seen == list(range(total))


def test_raw_sample_packing_roundtrip():
    from corintho_ai_b200.dist import pack_raw_samples, unpack_raw_samples
    rng = np.random.default_rng(0)
    st = rng.integers(0, 2**63, size=(11, 2), dtype=np.uint64)
    pr = rng.random((11, 96), dtype=np.float32)
    lb = np.array([1, -1, 0.0, -0.0] * 3, np.float32)[:11]
    go = np.arange(11, dtype=np.int32)
    s2, p2, l2, g2 = unpack_raw_samples(pack_raw_samples(st, pr, lb, go, first_game=5))
    assert (s2 == st).all() and p2.tobytes() == pr.tobytes() and l2.tobytes() == lb.tobytes()
    assert (g2 == go + 5).all()


def test_two_rank_gloo_gather_equals_single_run(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from oracle.pyoracle import OracleLib
    gs, pr, ev, r = identity_rows(OracleLib())
    want = np.concatenate([gs, pr, ev[:, None]], 1).astype(np.float32)
    for rank in range(2):
        got = np.load(tmp_path / f"rank{rank}.npy")
        assert got.tobytes() == want.tobytes()
        assert int(np.load(tmp_path / f"counts{rank}.npy").sum()) == r["num_samples"]

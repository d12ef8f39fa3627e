#!/usr/bin/env python
"""BASELINE.json configs[4]: evaluation match between two random-init networks -- 10 000 games,
1600 sims/move, no root noise, two-model testing mode (main.pyx:329-350, trainer.cpp:59-68) --
sharded over the ranks of one box (rank r owns a contiguous block of global game indices; game
parity = global index & 1, so model 0 moves first in even games).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 tools/match_bench.py [--games 10000] [--sims 1600]
    python tools/match_bench.py --games 1250          (one GPU, one shard's worth)

Prints one JSON line: games/s, sims/s, model-0 score over ALL games (Trainer::score arithmetic on
the all-gathered per-game results), and the result of the shard cross-check: rank 0 re-plays a
64-game slice that belongs to the LAST rank as its own shard and must get the same results.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def trainer_score(results):
    """Trainer::score (trainer.cpp:59-68) from per-game result codes, float32 like the reference:
    even games count the first player's score, odd games 1 - it (model 0 moves second there)."""
    gs = np.where(results == 1, 0.0, np.where(results == 3, 1.0, 0.5)).astype(np.float32)
    score = np.float32(0)
    for v in gs[0::2]:
        score = np.float32(score + v)
    for v in gs[1::2]:
        score = np.float32(np.float64(score) + (1.0 - np.float64(v)))
    return float(np.float32(score / np.float32(len(gs))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=10000)
    ap.add_argument("--sims", type=int, default=1600)
    ap.add_argument("--spe", type=int, default=16)
    ap.add_argument("--runs", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import corintho_ai_b200 as cb
    from corintho_ai_b200.dist import shard_range
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cb.lib().cb200_set_device(local)
    dev = torch.device("cuda", local)
    first, count = shard_range(args.games, world, rank)
    fa, fb = cb.fold_batchnorm(cb.random_weights(1)), cb.fold_batchnorm(cb.random_weights(2))

    def make(first_game, n):
        t = cb.Trainer(n, "", 12345, args.sims, args.spe, 1.0, 0.0, 0, 1, True,
                       total_games=args.games, first_game=first_game)
        t.set_weights(fa, 0, "bf16")
        t.set_weights(fb, 1, "bf16")
        return t

    t = make(first, count)
    t.run_selfplay(0)  # warm-up
    ms = []
    for k in range(args.runs):
        t.reset(12345)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t.run_selfplay(0)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    c = t.counters()
    res = torch.from_numpy(t.game_results().astype(np.int32)).to(dev)
    stats = torch.tensor([max(ms[1:] or ms), c["simulations"], c["moves"]], device=dev, dtype=torch.float64)
    all_res = res
    if dist is not None:
        # variable shard sizes: pad to the largest
        n_max = (args.games + world - 1) // world
        pad = torch.zeros(n_max, dtype=torch.int32, device=dev)
        pad[:count] = res
        out = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(out, pad)
        all_res = torch.cat([o[:shard_range(args.games, world, r)[1]] for r, o in enumerate(out)])
        tmax = stats[:1].clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tot = stats[1:].clone()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        stats = torch.cat([tmax, tot])
    all_res = all_res.cpu().numpy()
    check = None
    if rank == 0:
        # shard cross-check: 64 games from the END of the global range, re-played on rank 0
        n_chk = min(64, args.games)
        f_chk = (args.games - n_chk) & ~1
        u = make(f_chk, args.games - f_chk)
        u.run_selfplay(0)
        check = bool((u.game_results() == all_res[f_chk:]).all())
        step_s = float(stats[0]) * 1e-3
        line = {
            "metric": "evaluation_match_games_per_sec", "value": args.games / step_s, "unit": "games/s",
            "n_gpus": world, "ms_per_match": float(stats[0]), "runs_ms_rank0": ms,
            "sims_per_sec": float(stats[1]) / step_s, "moves_per_sec": float(stats[2]) / step_s,
            "model0_score": trainer_score(all_res), "games": args.games, "games_finished": int((all_res != 0).sum()),
            "results_histogram": {"first_player_lost": int((all_res == 1).sum()), "draw": int((all_res == 2).sum()),
                                  "first_player_won": int((all_res == 3).sum())},
            "shard_cross_check_last_%d_games_replayed_on_rank0" % (args.games - f_chk): check,
            "config": {"workload": "BASELINE.json configs[4]: %d games, %d sims/move, spe %d, epsilon 0, two random-init "
                                   "networks (bf16), testing mode, games sharded contiguously over %d GPU(s)"
                                   % (args.games, args.sims, args.spe, world)},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

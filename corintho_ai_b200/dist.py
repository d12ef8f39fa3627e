"""Multi-GPU plumbing for the self-play path: contiguous game shards per rank and the one
collective of the path -- an all-gather of the finished training samples.
torch.distributed is used as plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(total_games, world_size, rank):
    """Rank r owns global games [first, first+count): contiguous blocks in index order, the
    remainder spread over the first ranks, so concatenating the ranks' outputs in rank order
    reproduces the single-trainer (and reference) order of Trainer::writeSamples
    (corintho_ai/cpp/src/trainer.cpp:103-113)."""
    base, rem = divmod(int(total_games), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def pack_raw_samples(states, probs, labels, game_of, first_game=0):
    """[n, 2] u64 cstates + [n, 96] f32 + [n] f32 + [n] i32 -> one [n, 102] f32 row block
    (cstate words bit-cast to 4 floats, global game index bit-cast to 1 float)."""
    n = states.shape[0]
    row = np.empty((n, 4 + 96 + 1 + 1), np.float32)
    row[:, :4] = np.ascontiguousarray(states, np.uint64).view(np.float32).reshape(n, 4)
    row[:, 4:100] = probs
    row[:, 100] = labels
    row[:, 101] = (np.asarray(game_of, np.int32) + np.int32(first_game)).view(np.float32)
    return row


def unpack_raw_samples(rows):
    rows = np.ascontiguousarray(rows, np.float32)
    n = rows.shape[0]
    states = np.ascontiguousarray(rows[:, :4]).view(np.uint64).reshape(n, 2)
    return states, rows[:, 4:100].copy(), rows[:, 100].copy(), np.ascontiguousarray(rows[:, 101]).view(np.int32)


def all_gather_rows(dist, rows, device):
    """Variable-length all-gather: every rank contributes rows [n_r, w]; returns the
    concatenation in rank order as a numpy array (identical on every rank)."""
    import torch
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(rows, np.float32)).to(device)
    n_loc = torch.tensor([t.shape[0]], device=device, dtype=torch.int64)
    counts = [torch.zeros_like(n_loc) for _ in range(world)]
    dist.all_gather(counts, n_loc)
    counts = [int(c.item()) for c in counts]
    n_max = max(max(counts), 1)
    pad = torch.zeros((n_max, t.shape[1]), device=device, dtype=torch.float32)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)], 0), counts


class _DeviceArray:
    """Minimal __cuda_array_interface__ wrapper so torch can view engine-owned device memory."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def device_rows_as_tensor(ptr, n_rows, width, device):
    import torch
    if n_rows == 0 or not ptr:
        return torch.empty((0, width), dtype=torch.float32, device=device)
    return torch.as_tensor(_DeviceArray(ptr, (n_rows, width)), device=device)


def all_gather_rows_device(dist, rows, device):
    """Variable-length all-gather of a device tensor [n_r, w] over NCCL; returns (tensor
    [sum n_r, w] in rank order on the device, counts)."""
    import torch
    world = dist.get_world_size()
    n_loc = torch.tensor([rows.shape[0]], device=device, dtype=torch.int64)
    counts = torch.zeros(world, device=device, dtype=torch.int64)
    dist.all_gather_into_tensor(counts, n_loc)
    counts = [int(c) for c in counts.tolist()]
    n_max = max(max(counts), 1)
    pad = torch.zeros((n_max, rows.shape[1]), device=device, dtype=torch.float32)
    pad[:rows.shape[0]] = rows
    out = torch.empty((world * n_max, rows.shape[1]), device=device, dtype=torch.float32)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * n_max:r * n_max + c] for r, c in enumerate(counts)], 0), counts

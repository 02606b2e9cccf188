"""Utterance sharding across the GPUs of one box.

Utterances are independent, so the denoising loop needs no collective: each rank (one process
per GPU) holds a full weight replica, takes a length-balanced subset of the utterances, runs its
own reverse loop with Philox noise keyed by the *global* utterance id (results are independent of
the number of ranks), and the generated codes are exchanged with ONE all-gather at the end.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def partition(costs: list[int], world: int) -> list[list[int]]:
    """Greedy longest-first assignment of utterance indices to ``world`` ranks (cost ~ T^2 + T)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0] * world
    parts: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (loads[j], j))
        parts[r].append(i)
        loads[r] += costs[i]
    return [sorted(p) for p in parts]


def utterance_cost(t_txt: int, t_prom: int, t_resp: int, d_model: int = 1024) -> int:
    """GEMM work ~ 24 T d^2 and attention ~ 4 T^2 d per layer."""
    T = t_txt + t_prom + t_resp + 2
    return 24 * T * d_model * d_model + 4 * T * T * d_model


def generate_sharded(generate_fn, text_list, proms_list, resp_lens, n_levels: int = 8, group=None,
                     device=None, d_model: int = 1024, packed: bool = False, timing: dict | None = None):
    """Runs ``generate_fn(text_sub, proms_sub, resp_lens_sub, gids) -> [LongTensor (t'', n_levels)]``
    on this rank's shard and returns the codes of ALL utterances (global order) on every rank.

    Every rank is given the same full lists (token ids are tiny); one all-gather moves the codes:
    int16 tensor (n_max, T_max, n_levels) per rank.  Works on NCCL (GPU tensors) and gloo (CPU).

    ``packed=True`` returns ``(codes int16 (n, T_max, n_levels) in global utterance order, zero padded,
    resp_lens)`` instead of a list — one tensor for a single device->host copy or the EnCodec hand-off.
    ``timing`` (optional dict) receives CUDA events around this rank's local generation (``local_start`` /
    ``local_end``), so a caller can tell a slow rank from a slow collective.
    """
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    n = len(text_list)
    costs = [utterance_cost(len(t), len(p), r, d_model) for t, p, r in zip(text_list, proms_list, resp_lens)]
    parts = partition(costs, world)
    mine = parts[rank]
    if timing is not None and torch.cuda.is_available():
        timing["local_start"] = torch.cuda.Event(enable_timing=True)
        timing["local_end"] = torch.cuda.Event(enable_timing=True)
        timing["local_start"].record()
    local = generate_fn([text_list[i] for i in mine], [proms_list[i] for i in mine],
                        [resp_lens[i] for i in mine], mine) if mine else []
    if timing is not None and "local_end" in timing:
        timing["local_end"].record()
    if world == 1 and not packed:
        out = [None] * n
        for i, codes in zip(mine, local):
            out[i] = codes
        return out
    n_max = max(len(p) for p in parts)
    t_max = max(resp_lens) if resp_lens else 0
    if device is None:
        device = local[0].device if local else torch.device("cpu")
    if local and all(c.shape[0] == t_max for c in local):          # equal lengths: one stack, no per-utterance copies
        send = torch.zeros(n_max, t_max, n_levels, dtype=torch.int16, device=device)
        send[: len(local)] = torch.stack(local).to(torch.int16)
    else:
        send = torch.zeros(n_max, t_max, n_levels, dtype=torch.int16, device=device)
        for j, codes in enumerate(local):
            send[j, : codes.shape[0]] = codes.to(torch.int16)
    if world > 1:
        recv = torch.empty(world, n_max, t_max, n_levels, dtype=torch.int16, device=device)
        # int16 is not a NCCL/gloo datatype: ship the same bytes as uint8
        dist.all_gather_into_tensor(recv.view(torch.uint8).view(-1), send.view(torch.uint8).view(-1), group=group)
    else:
        recv = send.unsqueeze(0)
    if packed:
        order = torch.empty(n, dtype=torch.long)
        for r in range(world):
            for j, i in enumerate(parts[r]):
                order[i] = r * n_max + j
        return recv.view(world * n_max, t_max, n_levels)[order.to(recv.device)], list(resp_lens)
    out = [None] * n
    for r in range(world):
        for j, i in enumerate(parts[r]):
            out[i] = recv[r, j, : resp_lens[i]].long()
    return out

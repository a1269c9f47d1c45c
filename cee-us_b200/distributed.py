"""Multi-GPU plumbing of the planner: trajectories are sharded across ranks (one process per GPU); the only exchange
per iCEM iteration is an all-gather of each rank's k elite-candidate records (cost, global index, costs-per-model,
action sequence) -- a few KB, latency bound.  Every rank then runs the same merge + refit kernel, so ``mean_/std_``
and the elites stay bit-identical across ranks without a broadcast.

Precedent upstream for sharding trajectories over workers: ParallelGroundTruthModel.predict_n_steps
(mbrl/models/gt_par_model.py:91-128, np.array_split at :103-105).
"""
import torch
import torch.distributed as dist


def shard_range(total, world_size, rank):
    """Contiguous shard [lo, hi) of ``total`` trajectories for ``rank`` (equal shards are required by the kernels)."""
    if total % world_size != 0:
        raise ValueError(f"num_simulated_trajectories={total} must be divisible by world_size={world_size}")
    per = total // world_size
    return rank * per, (rank + 1) * per


def gather_candidates(cand, group=None):
    """cand [nE,k,rec] (this rank) -> cand_all [world,nE,k,rec], identical on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return cand.unsqueeze(0)
    world = dist.get_world_size(group)
    out = torch.empty((world * cand.shape[0],) + tuple(cand.shape[1:]), dtype=cand.dtype, device=cand.device)
    dist.all_gather_into_tensor(out, cand.contiguous(), group=group)  # rank-major concatenation along dim 0
    return out.view((world,) + tuple(cand.shape))


def merge_candidates_reference(cand_all, k):
    """Host-side statement of what the merge kernel does with the gathered records: global top-k by (cost, index).
    Used by the CPU (gloo) tests of the sharding logic; the product path runs update_kernel on the GPU."""
    flat = cand_all.reshape(-1, cand_all.shape[-1])
    cost = flat[:, 0]
    idx = flat[:, 1].contiguous().view(torch.int32).to(torch.int64)
    order = sorted(range(flat.shape[0]), key=lambda i: (float(cost[i]), int(idx[i])))[:k]
    return flat[order], idx[order]

"""Multi-GPU plumbing: one process per GPU (torch.distributed), independent proofs sharded by contiguous
blocks, one broadcast of the generator table at start-up, verdicts gathered on rank 0.  No data-path
collective: a proof is verified entirely on the rank that owns it."""
import torch
import torch.distributed as dist


def shard(count: int, rank: int, world: int):
    """[start, stop) of the contiguous block of `count` independent items owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(count, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def broadcast_table(table: torch.Tensor, src: int = 0):
    """The single collective of the path: the generator table (bp_gens_device_table bytes) from rank `src`."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(table, src=src)
    return table


def gather_verdicts(local: bytes, count: int, device="cpu"):
    """Per-proof verdict codes of every rank's shard, concatenated in proof order on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return bytes(local)
    world = dist.get_world_size()
    sizes = [shard(count, r, world)[1] - shard(count, r, world)[0] for r in range(world)]
    width = max(sizes) if sizes else 0
    mine = torch.zeros(max(width, 1), dtype=torch.uint8, device=device)
    if len(local):
        mine[:len(local)] = torch.frombuffer(bytearray(local), dtype=torch.uint8).to(device)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return b"".join(bytes(p[:s].cpu().numpy().tobytes()) for p, s in zip(parts, sizes))


def verify_sharded(verify_fn, proofs: bytes, commitments: bytes, proof_len: int, m: int, count: int, device="cpu"):
    """Verify `count` proofs across the ranks: rank r runs verify_fn(proofs_shard, commitments_shard, n_shard) -> verdict
    bytes on its block and every rank returns the full verdict list.  verify_fn is bp.verify_batch bound to the rank's
    context on a GPU box (tests inject the oracle to exercise the plumbing on CPU with gloo)."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    a, b = shard(count, rank, world)
    local = verify_fn(proofs[a * proof_len:b * proof_len], commitments[a * 32 * m:b * 32 * m], b - a) if b > a else b""
    return list(gather_verdicts(bytes(local), count, device=device))

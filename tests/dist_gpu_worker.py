"""Worker of tests/test_gpu_dist.py: ONE batch of range proofs sharded across the ranks (north_star: "independent proofs in a batch
shard one-per-GPU"), each shard verified by bp.verify_batch on the rank's GPU, verdicts gathered and compared with the oracle.
Launched by torchrun (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* in the environment)."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    ngpu = torch.cuda.device_count()
    one_gpu_each = ngpu >= world
    dev = local if one_gpu_each else 0                      # a 1-GPU box: both ranks share cuda:0, the plumbing runs over gloo
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if one_gpu_each else "gloo", device_id=torch.device("cuda", dev) if one_gpu_each else None)
    import bulletproofs_b200 as bp
    from bulletproofs_b200.dist import broadcast_table, verify_sharded
    from oracle_binding import Oracle, L_ORDER
    orc = Oracle()
    n, m, count = 32, 1, 45                                # odd count: shards of different sizes
    label = b"sharded batch"
    og = orc.gens(n, m); ot = orc.transcript(label)
    rnd = random.Random(77)
    values = [rnd.randrange(1 << n) for _ in range(count)]
    blind = b"".join(rnd.randrange(L_ORDER).to_bytes(32, "little") for _ in range(count))
    seeds = b"".join(i.to_bytes(8, "little") + bytes(24) for i in range(count))
    proofs, Vs = orc.prove_many(og, ot, values, blind, n, m, seeds, nthreads=2)            # identical on every rank
    plen = len(proofs) // count
    pb = bytearray(proofs)
    damaged = [0, 13, 22, 23, 44]                          # both shards, shard boundary, first and last proof
    for i in damaged:
        pb[i * plen + 100] ^= 4
    want = orc.verify_many(og, ot, bytes(pb), plen, Vs, n, m, count, nthreads=2)
    ctx = bp.Context(dev)
    # the generator table: derived on rank 0 only, broadcast, imported on the others
    gens = bp.Gens(ctx, n, m, empty=(rank != 0))
    _, nbytes = gens.device_table()
    table = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    if rank == 0:
        gens.table_export(table.data_ptr())
    if one_gpu_each:
        broadcast_table(table)
    else:
        host = table.cpu(); broadcast_table(host); table.copy_(host)
    torch.cuda.synchronize()
    if rank != 0:
        gens.table_import(table.data_ptr())
    assert gens.G(0, 3) == orc.gens_get(og, 0, 0, 3)
    t = bp.Transcript(label)
    calls = []

    def verify_fn(p, v, k):
        calls.append(k)
        return bytes(bp.verify_batch(ctx, gens, t, p, v, n, m, k))

    got = verify_sharded(verify_fn, bytes(pb), Vs, plen, m, count, device="cuda" if one_gpu_each else "cpu")
    assert got == want, (rank, [(i, g, w) for i, (g, w) in enumerate(zip(got, want)) if g != w])
    assert [i for i, v in enumerate(got) if v] == damaged
    assert calls and calls[0] in (count // world, count // world + 1) and ctx.launches > 0
    print(f"rank {rank}/{world} on cuda:{dev} ({'nccl' if one_gpu_each else 'gloo'}): shard of {calls[0]} proofs, {ctx.launches} launches, verdicts ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

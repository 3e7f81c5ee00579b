"""CPU tier: the N>1 host logic on two gloo ranks — shard partition, the table broadcast, verdict gathering.
The per-shard verifier is the oracle here (no GPU in this tier); on a GPU box the same plumbing wraps bp.verify_batch."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_partition():
    from bulletproofs_b200.dist import shard
    for count in (0, 1, 2, 7, 8, 1023, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard(count, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == count
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, count, out_dir):
    import sys
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bulletproofs_b200.dist import broadcast_table, verify_sharded
    from oracle_binding import Oracle, L_ORDER
    orc = Oracle()
    # (1) the one collective: rank 0's table reaches every rank unchanged
    g = torch.Generator().manual_seed(5)
    table = torch.randint(0, 256, (96 * 130,), dtype=torch.uint8, generator=g) if rank == 0 else torch.zeros(96 * 130, dtype=torch.uint8)
    broadcast_table(table)
    ref = torch.randint(0, 256, (96 * 130,), dtype=torch.uint8, generator=torch.Generator().manual_seed(5))
    assert torch.equal(table, ref)
    # (2) shard + verify + gather: identical workload on every rank, damaged proofs found at the right global positions
    n, m = 8, 1
    og = orc.gens(8, 1); label = b"dist test"; t = orc.transcript(label)
    rnd = random.Random(3)
    values = [rnd.randrange(1 << n) for _ in range(count)]
    blind = b"".join(rnd.randrange(L_ORDER).to_bytes(32, "little") for _ in range(count))
    seeds = b"".join(i.to_bytes(8, "little") + bytes(24) for i in range(count))
    proofs, Vs = orc.prove_many(og, t, values, blind, n, m, seeds, nthreads=2)
    plen = len(proofs) // count
    pb = bytearray(proofs)
    damaged = [1, count // 2, count - 1]
    for i in damaged:
        pb[i * plen + 40] ^= 1
    calls = []

    def verify_fn(p, v, k):
        calls.append(k)
        return bytes(orc.verify_many(og, t, p, plen, v, n, m, k, nthreads=2))

    verdicts = verify_sharded(verify_fn, bytes(pb), Vs, plen, m, count)
    assert [i for i, v in enumerate(verdicts) if v] == damaged
    assert len(verdicts) == count and sum(calls) in (count // world, count - count // world, (count + world - 1) // world)
    with open(os.path.join(out_dir, f"rank{rank}.ok"), "w") as f:
        f.write(",".join(map(str, verdicts)))
    dist.destroy_process_group()


@pytest.mark.parametrize("count", [9, 16])
def test_two_rank_gloo(built, tmp_path, count):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, count, str(tmp_path)), nprocs=2, join=True)
    a = open(tmp_path / "rank0.ok").read(); b = open(tmp_path / "rank1.ok").read()
    assert a == b and len(a.split(",")) == count

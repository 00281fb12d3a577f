"""N>1 host logic without a GPU: the stream -> rank shard map and the barrier / max-over-ranks timing reduction bench.py uses,
run as two gloo ranks on the CPU (streams are sharded, there is no data-path collective)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_map_partitions_the_streams():
    import bench
    for world in (1, 2, 4, 8):
        parts = [bench.shard(1024, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(1024))          # every stream on exactly one rank
        assert {len(p) for p in parts} == {1024 // world}                         # balanced
        assert all(i % world == r for r, p in enumerate(parts) for i in p)        # stream i -> rank i mod N
    parts = [bench.shard(10, r, 4) for r in range(4)]                             # ragged
    assert [len(p) for p in parts] == [3, 3, 2, 2]
    assert bench.shard(0, 0, 2) == []                                             # empty


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    mine = bench.shard(64, rank, world)
    # each rank "processes" its shard; timings differ per rank; the job-level figure uses the MAX over ranks
    local_s = 0.010 * (rank + 1)
    agg = bench.reduce_timings([local_s, 2 * local_s, 3 * local_s], world, device="cpu")
    n = torch.tensor([len(mine)], dtype=torch.int64)
    dist.all_reduce(n)                      # test-only: total units over ranks
    dist.barrier()
    if rank == 0:
        torch.save({"agg": agg, "units": int(n)}, out)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_timing_reduction_gloo(tmp_path):
    out = str(tmp_path / "r.pt")
    mp.spawn(_rank_main, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["units"] == 64
    assert r["agg"] == pytest.approx([0.020, 0.040, 0.060])       # max over the two ranks

"""The N>1 path on CPU: world_size-2 gloo processes exercise the only cross-rank steps the bench has (barrier, MAX of the
timed region, SUM of the Report counters) and the unit sharding.  No data-path collective exists (pairs are independent)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import tidalwave_b200 as tw
    d = tw.dist.Dist(backend="gloo")
    assert d.world == 2
    mine = list(tw.dist.shard(11, d.rank, d.world))
    d.barrier()
    mx = d.reduce_max(10.0 + d.rank * 5)          # rank 1 is slower
    report = d.reduce_sum([len(mine), len(mine) - d.rank, d.rank])  # request, data, error
    val = tw.dist.whole_job_throughput(len(mine), d.world, mx)
    print(json.dumps(dict(rank=d.rank, mine=mine, mx=mx, report=report, val=val)), flush=True)
    d.close()
""") % ROOT


def test_two_rank_gloo(tmp_path):
    port = _free_port()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0, err[-2000:]
        outs.append(__import__("json").loads(out.strip().splitlines()[-1]))
    outs.sort(key=lambda o: o["rank"])
    assert sorted(outs[0]["mine"] + outs[1]["mine"]) == list(range(11))     # every pair on exactly one rank
    assert outs[0]["mx"] == outs[1]["mx"] == 15.0                          # max over ranks
    assert outs[0]["report"] == outs[1]["report"] == [11.0, 10.0, 1.0]     # Report counters summed
    assert abs(outs[0]["val"] - 6 * 2 / 15.0) < 1e-12


def test_shard_covers_everything():
    import tidalwave_b200 as tw
    for n in (0, 1, 7, 64, 10000):
        for world in (1, 2, 4, 8):
            got = [i for r in range(world) for i in tw.dist.shard(n, r, world)]
            assert got == list(range(n))
            sizes = [len(tw.dist.shard(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1

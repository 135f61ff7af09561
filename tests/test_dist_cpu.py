"""CPU tests of the multi-rank host logic with the gloo backend (world_size 2)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_panel_ownership_is_a_partition(pkg):
    d = pkg.dist
    for n, world in ((40000, 8), (5000, 2), (300, 4), (64, 1)):
        pw = d.panel_width(n)
        assert pw % 64 == 0
        owners = [d.panel_owner(c, pw, world) for c in range(n)]
        assert set(owners) <= set(range(world))
        # block-cyclic: constant inside a panel, rotates between consecutive panels
        for c in range(0, n - pw, pw):
            assert len(set(owners[c:c + pw])) == 1
            if world > 1:
                assert owners[c + pw] == (owners[c] + 1) % world
        counts = [owners.count(r) for r in range(world)]
        assert max(counts) - min(counts) <= pw


def test_unique_id_exchange_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as dist
        import __graft_entry__ as g
        pkg = g.load_package()
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        uid = pkg.dist.exchange_unique_id(lambda: bytes(range(128)), rank, world)
        assert uid == bytes(range(128)), uid
        # the bench's reduction of per-rank timings (max over ranks)
        import torch
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == float(world)
        open(os.path.join({str(tmp_path)!r}, f"ok{{rank}}"), "w").write("ok")
        dist.destroy_process_group()
    """))
    port = _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()

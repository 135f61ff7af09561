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


def test_row_ownership_is_a_partition(pkg):
    d = pkg.dist
    for n, world in ((40000, 8), (5000, 2), (300, 4), (64, 1)):
        pw = d.block_rows(n)
        assert pw % 64 == 0
        owners = [d.row_owner(r, pw, world) for r in range(n)]
        assert set(owners) <= set(range(world))
        # block-cyclic: constant inside a row block, rotates between consecutive blocks
        for r in range(0, n - pw, pw):
            assert len(set(owners[r:r + pw])) == 1
            if world > 1:
                assert owners[r + pw] == (owners[r] + 1) % world
        counts = [owners.count(r) for r in range(world)]
        assert max(counts) - min(counts) <= pw
        nblk = -(-n // pw)
        for p in range(nblk):
            # the slots of the all-gather of step p cover every row block >= p exactly once
            blocks = []
            for r in range(world):
                f, c = d.first_block(p, r, world), d.count_blocks(p, r, world, nblk)
                blocks += [f + z * world for z in range(c)]
                assert c <= -(-(nblk - p) // world)
            assert sorted(blocks) == list(range(p, nblk))


def _run_world2(tmp_path, body):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(body))
    port = _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-3000:]
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_unique_id_exchange_gloo_world2(tmp_path):
    _run_world2(tmp_path, f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import torch.distributed as dist
        import __graft_entry__ as g
        pkg = g.load_package()
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        uid = pkg.dist.exchange_unique_id(lambda: bytes(range(128)), rank, world)
        assert uid == bytes(range(128)), uid
        # the bench's reduction of per-rank timings (max over ranks) and its agreed restart decision
        import torch
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == float(world)
        open(os.path.join({str(tmp_path)!r}, f"ok{{rank}}"), "w").write("ok")
        dist.destroy_process_group()
    """)


def test_row_block_cyclic_cholesky_schedule_gloo_world2(tmp_path):
    """The communication schedule of csrc/dist.cu (cholesky_dist) replayed with NumPy blocks over gloo, world 2: every rank
    holds only its own row blocks of the lower triangle; per step the owner factors the diagonal block and broadcasts the
    inverse of its factor, every rank solves its own rows of the panel, the solved blocks travel in ONE all-gather whose slot
    layout comes from dist.first_block / dist.count_blocks, and every rank updates its own row blocks up to their diagonal.
    All ranks must end with the complete factor.  Sizes include a partial last block."""
    _run_world2(tmp_path, f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        import __graft_entry__ as g
        pkg = g.load_package()
        d = pkg.dist
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        for n, pw in ((100, 16), (96, 32), (70, 64)):
            rng = np.random.default_rng(n)
            M = rng.standard_normal((n, n)); H = M @ M.T + n * np.eye(n)
            nblk = -(-n // pw)
            A = np.full((n, n), np.nan)                          # only my row blocks of the lower triangle are valid
            for gblk in range(rank, nblk, world):
                r0, r1 = gblk * pw, min(n, (gblk + 1) * pw)
                A[r0:r1, :r1] = np.tril(H)[r0:r1, :r1]
            for p in range(nblk):
                c0, c1 = p * pw, min(n, (p + 1) * pw); w = c1 - c0; owner = p % world
                X = torch.zeros(pw, pw, dtype=torch.float64)
                if rank == owner:
                    Lpp = np.linalg.cholesky(np.tril(A[c0:c1, c0:c1]) + np.tril(A[c0:c1, c0:c1], -1).T)
                    A[c0:c1, c0:c1] = Lpp
                    X[:w, :w] = torch.from_numpy(np.linalg.inv(Lpp))
                dist.broadcast(X, src=owner)
                Xn = X.numpy()[:w, :w]
                maxcnt = -(-(nblk - p) // world)
                send = torch.zeros(maxcnt, pw, pw, dtype=torch.float64)
                f, cnt = d.first_block(p, rank, world), d.count_blocks(p, rank, world, nblk)
                for z in range(cnt):
                    gblk = f + z * world; r0, r1 = gblk * pw, min(n, (gblk + 1) * pw)
                    blk = A[c0:c1, c0:c1] if gblk == p else A[r0:r1, c0:c1] @ Xn.T
                    send[z, :r1 - r0, :w] = torch.from_numpy(np.ascontiguousarray(blk))
                recv = [torch.zeros_like(send) for _ in range(world)]
                dist.all_gather(recv, send)
                for r in range(world):
                    fr, cr = d.first_block(p, r, world), d.count_blocks(p, r, world, nblk)
                    for z in range(cr):
                        gblk = fr + z * world; r0, r1 = gblk * pw, min(n, (gblk + 1) * pw)
                        A[r0:r1, c0:c1] = recv[r][z, :r1 - r0, :w].numpy()
                for gblk in range(d.first_block(p + 1, rank, world), nblk, world):
                    r0, r1 = gblk * pw, min(n, (gblk + 1) * pw)
                    A[r0:r1, c1:r1] -= A[r0:r1, c0:c1] @ A[c1:r1, c0:c1].T
            Lref = np.linalg.cholesky(H)
            err = np.linalg.norm(np.tril(A) - Lref) / np.linalg.norm(Lref)
            assert err < 1e-12, (n, pw, err)
        open(os.path.join({str(tmp_path)!r}, f"ok{{rank}}"), "w").write("ok")
        dist.destroy_process_group()
    """)

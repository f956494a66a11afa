"""The N>1 path on CPU: two gloo ranks, each rendering its interleaved row-block tile with the oracle, must
reassemble the single-rank image bit for bit (the reference's seed depends only on global pixel coordinates and the
global sample index, raygen.rgen:47-48), and bench.py's rank plumbing (barrier, max/sum over ranks, byte broadcast,
tile assignment) must behave. No GPU, no libbpt compute calls."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle_lib as O  # noqa: E402

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import bench, oracle_lib as O
d = bench.Dist(2, backend="gloo")
assert d.active and d.world == 2
tile = bench.tile_kwargs(64, d.world, d.rank)
assert tile == dict(tile_block=8, tile_nranks=2, tile_rank=d.rank)
verts, idx, faces, _ = O.load_cornell_golden()
scene = O.Scene(verts, idx, faces)
img = np.zeros((64, 48, 4), np.float32)
rays = 0
for f in range(2):
    _, r = scene.render(O.default_params(48, 64, 2, 4, f, **tile), 32, nthreads=1, image=img)
    rays += r
mine = bench.interleaved_rows(64, 8, 2, d.rank)
other = bench.interleaved_rows(64, 8, 2, 1 - d.rank)
assert np.all(img[other] == 0) and np.all(img[mine, :, 3] == 1)
import torch
t = torch.from_numpy(img)
d.dist.all_reduce(t)                       # tiles are disjoint: the sum is the assembled image
total = d.sum(rays)
assert d.max(d.rank + 1) == 2.0
blob = d.broadcast_bytes(bytes(range(128)) if d.rank == 0 else b"", 128)
assert blob == bytes(range(128))
d.barrier()
if d.rank == 0:
    np.save({out!r}, t.numpy())
    open({out!r} + ".rays", "w").write(str(int(total)))
d.close()
"""


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_interleaved_rows_partition():
    for (h, b, n) in ((64, 8, 2), (4096, 8, 8), (1080, 1, 8), (96, 4, 3)):
        rows = [bench.interleaved_rows(h, b, n, r) for r in range(n)]
        assert sorted(sum(rows, [])) == list(range(h))
        assert len({len(r) for r in rows}) == 1
    assert bench.tile_kwargs(4096, 1, 0) == {}
    assert bench.tile_kwargs(4096, 8, 3) == dict(tile_block=8, tile_nranks=8, tile_rank=3)
    assert bench.tile_kwargs(1080, 2, 1)["tile_block"] == 4   # 1080 = 2^3 * 135: largest power-of-two block dividing 540


def test_two_gloo_ranks_reassemble_the_image(tmp_path, cornell_oracle):
    out = str(tmp_path / "img.npy")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=out))
    port = free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        log, _ = p.communicate(timeout=300)
        assert p.returncode == 0, log
    got = np.load(out)
    ref = np.zeros((64, 48, 4), np.float32)
    rays = 0
    for f in range(2):
        _, r = cornell_oracle.render(O.default_params(48, 64, 2, 4, f), 32, image=ref)
        rays += r
    assert np.array_equal(got, ref)
    assert int(open(out + ".rays").read()) == rays


def test_reference_arm_prints_the_contract_line():
    """--impl reference on a small Cornell workload: one JSON line with the keys the driver reads."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cornell",
                        "--width", "64", "--height", "64", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert len(r.stdout.strip().splitlines()) == 1, r.stdout   # stdout carries the JSON line and nothing else
    line = json.loads(r.stdout.strip())
    assert line["impl"] == "reference" and line["metric"] == "Mray/s" and line["value"] > 0
    # the reference's shader text compiled as C++ where oracle/_ref/libref_shade.so exists, else the oracle's restatement
    assert line["cpu_baseline"]["kind"] == ("reference" if O.ref_shade_available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    # rank != 0 of a torchrun launch prints nothing and exits 0
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=dict(os.environ, RANK="1", WORLD_SIZE="2"), capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == ""

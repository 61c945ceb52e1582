"""The svo_builder / svo_builder_binary executables: CPU tests for argument handling
(the reference's quirks, main.cpp:100-196), GPU tests for byte-identical output files."""
import os
import subprocess

import pytest

from ooc_svo_builder_b200 import meshgen as mg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "ooc_svo_builder_b200", "bin")


def run(exe, *args):
    p = subprocess.run([os.path.join(BIN, exe), *args], capture_output=True, text=True, timeout=600)
    return p.returncode, p.stdout


def test_help_and_errors_exit_zero(tmp_path):
    for exe in ("svo_builder", "svo_builder_binary"):
        assert os.access(os.path.join(BIN, exe), os.X_OK)
        rc, out = run(exe, "-h", "x")
        assert rc == 0 and "-f <filename.tri>" in out
        rc, out = run(exe)
        assert rc == 0 and "Not enough or invalid arguments" in out               # main.cpp:104-107
        rc, out = run(exe, "-f", "mesh.obj")
        assert rc == 0 and "does not end in .tri" in out                          # main.cpp:112-117
        rc, out = run(exe, "-f", "a.tri", "-s", "100")
        assert rc == 0 and "not a power of 2" in out                              # main.cpp:121-126
        rc, out = run(exe, "-f", "a.tri", "-l", "1")
        assert rc == 0 and "nonsensical" in out                                   # main.cpp:130-135
        rc, out = run(exe, "-f", "a.tri", "-x")
        assert rc == 0 and "invalid arguments" in out                             # main.cpp:183-185
        rc, out = run(exe, "-f", str(tmp_path / "missing.tri"))
        assert rc == 0 and "does not exist" in out


def test_wrong_executable_for_tri_kind(tmp_path):
    geo = mg.write_tri(str(tmp_path / "geo"), mg.icosphere(1))
    pay = mg.write_tri(str(tmp_path / "pay"), mg.Mesh(mg.with_payload(mg.icosphere(1).tris), 2.0))
    rc, out = run("svo_builder", "-f", geo)
    assert rc == 0 and "contains only geometry" in out                            # main.cpp:261-265
    rc, out = run("svo_builder_binary", "-f", pay)
    assert rc == 0 and "more than just geometry" in out                           # main.cpp:256-260
    rc, out = run("svo_builder_binary", "-f", geo, "-c", "linear", "-h")
    assert "only doing binary voxelisation" in out                                # main.cpp:156-157


@pytest.mark.gpu
@pytest.mark.parametrize("payload,args", [(False, ["-s", "128"]), (False, ["-s", "256", "-l", "3", "-d", "20"]),
                                          (True, ["-s", "128"]), (True, ["-s", "128", "-c", "normal"]),
                                          (True, ["-s", "64", "-levels"])])
def test_cli_output_files_match_oracle(tmp_path, oracle, payload, args):
    m = mg.icosphere(4)
    if payload:
        m = mg.Mesh(mg.with_payload(m.tris), m.length)
    hdr = mg.write_tri(str(tmp_path / "mesh"), m)
    exe = "svo_builder" if payload else "svo_builder_binary"
    rc, out = run(exe, "-f", hdr, *args, "-v")
    assert rc == 0, out
    g = int(args[args.index("-s") + 1])
    lim = int(args[args.index("-l") + 1]) if "-l" in args else 2048
    color = args[args.index("-c") + 1] if "-c" in args else "model"
    want = oracle.build(m.tris, m.length, g, memory_limit_mb=lim, levels="-levels" in args, color=color)
    base = str(tmp_path / "mesh") + "%d_%d" % (g, want.n_partitions)
    got = oracle.read_outputs(base)
    assert got.header == want.header, out
    assert got.nodes == want.nodes
    assert got.data == want.data
    assert ("Total amount of voxels: %d" % want.n_voxels) in out
    if oracle.ref_available():      # and against the real reference binary when it travelled with the repo
        ref = oracle.ref_build(m, g, memory_limit_mb=lim if "-l" in args else None, levels="-levels" in args,
                               color=color if "-c" in args else None)
        assert (ref.header, ref.nodes, ref.data) == (got.header, got.nodes, got.data)


@pytest.mark.gpu
def test_cli_threaded_io_with_a_small_budget(tmp_path, oracle):
    """-l 2: the input ring and the output chunks shrink to a few hundred KB, so the file is read by several pread
    threads in many chunks and written by several pwrite threads: same bytes, and the -v lines of the reference."""
    m = mg.displaced_sphere(120, 120, seed=8)
    hdr = mg.write_tri(str(tmp_path / "mesh"), m)
    rc, out = run("svo_builder_binary", "-f", hdr, "-s", "256", "-l", "2", "-v")
    assert rc == 0, out
    want = oracle.build(m.tris, m.length, 256, memory_limit_mb=2)
    got = oracle.read_outputs(str(tmp_path / "mesh") + "256_%d" % want.n_partitions)
    assert (got.header, got.nodes, got.data) == (want.header, want.nodes, want.data), out[-1500:]
    found = [int(line.split()[1]) for line in out.splitlines() if line.strip().startswith("found ")]
    assert found and sum(found) == want.n_voxels                      # main.cpp:348, per partition
    assert "Total amount of voxels: %d" % want.n_voxels in out


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [2, 8])
@pytest.mark.parametrize("payload", [False, True])
def test_cli_multi_gpu_files_match_oracle(tmp_path, oracle, gpus, payload):
    import torch
    if torch.cuda.device_count() < gpus:
        pytest.skip("needs %d GPUs, this box has %d" % (gpus, torch.cuda.device_count()))
    m = mg.displaced_sphere(150, 150, seed=9)
    if payload:
        m = mg.Mesh(mg.with_payload(m.tris), m.length)
    hdr = mg.write_tri(str(tmp_path / "mesh"), m)
    exe = "svo_builder" if payload else "svo_builder_binary"
    rc, out = run(exe, "-f", hdr, "-s", "512", "-l", "100", "-gpus", str(gpus), "-v")
    assert rc == 0, out
    want = oracle.build(m.tris, m.length, 512, memory_limit_mb=100)
    got = oracle.read_outputs(str(tmp_path / "mesh") + "512_%d" % want.n_partitions)
    assert (got.header, got.nodes, got.data) == (want.header, want.nodes, want.data), out[-1500:]


@pytest.mark.gpu
@pytest.mark.parametrize("payload", [False, True])
def test_cli_multi_gpu_levels_files_match_oracle(tmp_path, oracle, payload):
    """-levels on the sharded path (internal nodes carry averaged data records; the ranks own contiguous ranges of BOTH files)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs, this box has %d" % torch.cuda.device_count())
    m = mg.displaced_sphere(120, 120, seed=11)
    if payload:
        m = mg.Mesh(mg.with_payload(m.tris), m.length)
    hdr = mg.write_tri(str(tmp_path / "mesh"), m)
    exe = "svo_builder" if payload else "svo_builder_binary"
    rc, out = run(exe, "-f", hdr, "-s", "256", "-l", "10", "-levels", "-gpus", "2")
    assert rc == 0, out
    want = oracle.build(m.tris, m.length, 256, memory_limit_mb=10, levels=True)
    got = oracle.read_outputs(str(tmp_path / "mesh") + "256_%d" % want.n_partitions)
    assert (got.header, got.nodes, got.data) == (want.header, want.nodes, want.data), out[-1500:]

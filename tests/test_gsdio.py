"""Minimal GSD reader/writer (cav_hoomd_b200/gsdio.py).  PARITY UNPINNED against the real
gsd library (absent from this image, and the reference ships no .gsd file): the reader is checked against
(a) a file assembled here byte by byte from the published file-layer specification, independently of the
writer, in both name-list layouts (file layer 1.0 and 2.x), and (b) writer -> reader round trips; the
bridge to the hot path's array layouts is checked against synth.make_system."""
import struct

import numpy as np
import pytest

from cav_hoomd_b200 import gsdio, synth


def _assemble(path, layer_major):
    """Two frames, N = 3; frame 1 carries only step and position (everything else falls back to frame 0)."""
    chunks = []  # (frame, name, N, M, type code, bytes)
    pos0 = np.array([[0.5, -1.25, 2.0], [3.0, 4.0, -5.0], [0.0, 0.0, 0.0]], dtype="<f4")
    pos1 = pos0 + np.float32(0.125)
    chunks.append((0, "configuration/step", 1, 1, 4, struct.pack("<Q", 7)))
    chunks.append((0, "configuration/box", 6, 1, 9, struct.pack("<6f", 10.0, 12.0, 14.0, 0.0, 0.0, 0.0)))
    chunks.append((0, "particles/N", 1, 1, 3, struct.pack("<I", 3)))
    chunks.append((0, "particles/types", 3, 2, 5, b"O\0N\0L\0"))
    chunks.append((0, "particles/typeid", 3, 1, 3, struct.pack("<3I", 0, 1, 2)))
    chunks.append((0, "particles/charge", 3, 1, 9, struct.pack("<3f", 0.5, -0.5, 0.0)))
    chunks.append((0, "particles/position", 3, 3, 9, pos0.tobytes()))
    chunks.append((0, "particles/image", 3, 3, 7, struct.pack("<9i", 1, 0, -1, 0, 0, 0, 2, -2, 0)))
    chunks.append((1, "configuration/step", 1, 1, 4, struct.pack("<Q", 1007)))
    chunks.append((1, "particles/position", 3, 3, 9, pos1.tobytes()))
    names = []
    for c in chunks:
        if c[1] not in names:
            names.append(c[1])
    body = b""
    index = b""
    for fr, name, n, m, code, raw in chunks:
        loc = 256 + len(body)
        body += raw
        index += struct.pack("<QQqIHBB", fr, n, loc, m, names.index(name), code, 0)
    index += b"\0" * 32 * 3  # unused entries (location 0)
    if layer_major == 1:
        nl = b"".join(n.encode().ljust(64, b"\0") for n in names) + b"\0" * 64
    else:
        nl = b"".join(n.encode() + b"\0" for n in names)
        nl = nl.ljust((len(nl) // 64 + 1) * 64, b"\0")
    nloc = 256 + len(body)
    iloc = nloc + len(nl)
    hdr = struct.pack("<QQQQQII64s64s80s", 0x65DF65DF65DF65DF, iloc, len(index) // 32, nloc, len(nl) // 64, (1 << 16) | 4,
                      layer_major << 16, b"test".ljust(64, b"\0"), b"hoomd".ljust(64, b"\0"), b"\0" * 80)
    assert len(hdr) == 256
    with open(path, "wb") as fh:
        fh.write(hdr + body + nl + index)
    return pos0, pos1


@pytest.mark.parametrize("layer", [1, 2])
def test_reader_on_a_file_assembled_from_the_specification(tmp_path, layer):
    path = str(tmp_path / "spec.gsd")
    pos0, pos1 = _assemble(path, layer)
    with gsdio.open_gsd(path) as f:
        assert len(f) == 2 and f.schema == "hoomd" and f.application == "test"
        a, b = f[0], f[-1]
        assert (a.step, b.step) == (7, 1007) and a.N == b.N == 3
        assert a.types == ["O", "N", "L"] and b.types == ["O", "N", "L"]
        assert np.array_equal(a.position, pos0) and np.array_equal(b.position, pos1)
        assert np.array_equal(b.typeid, [0, 1, 2]) and np.array_equal(b.charge, np.float32([0.5, -0.5, 0.0]))  # frame 0's
        assert np.array_equal(a.image, [[1, 0, -1], [0, 0, 0], [2, -2, 0]]) and np.array_equal(b.image, a.image)
        assert np.array_equal(a.mass, np.ones(3, np.float32)) and np.array_equal(a.velocity, np.zeros((3, 3), np.float32))  # defaults
        assert np.array_equal(a.box, np.float32([10, 12, 14, 0, 0, 0])) and a.dimensions == 3
        with pytest.raises(IndexError):
            f[2]
        assert f.chunk(1, "particles/charge") is None and f.chunk(0, "no/such/chunk") is None


def test_bad_files_are_rejected(tmp_path):
    p = tmp_path / "x.gsd"
    p.write_bytes(b"\0" * 100)
    with pytest.raises(ValueError):
        gsdio.open_gsd(str(p))
    p.write_bytes(b"\1" * 300)
    with pytest.raises(ValueError):
        gsdio.open_gsd(str(p))


@pytest.mark.parametrize("layer", [(1, 0), (2, 0)])
def test_write_read_round_trip_and_header_layout(tmp_path, layer):
    s = synth.make_system(1000, replica=3)
    fr0 = gsdio.system_to_frame(s, step=0)
    fr1 = gsdio.system_to_frame(s, step=500)
    fr1.position = fr1.position + np.float32(0.5)
    path = str(tmp_path / "traj.gsd")
    gsdio.write_gsd(path, [fr0, fr1], file_layer=layer)
    raw = open(path, "rb").read()
    magic, iloc, ialloc, nloc, nalloc, sver, gver = struct.unpack_from("<QQQQQII", raw)
    assert magic == 0x65DF65DF65DF65DF and gver == (layer[0] << 16) and sver == (1 << 16) | 4
    assert raw[112:117] == b"hoomd" and raw[48:55] == b"cavb200" and len(raw) == iloc + 32 * ialloc          # index is the last block
    assert raw[iloc + 32 * (ialloc - 1):] == b"\0" * 32                         # terminated by an unused entry
    with gsdio.open_gsd(path) as f:
        assert len(f) == 2
        for got, want in ((f[0], fr0), (f[1], fr1)):
            assert got.step == want.step and got.N == want.N and got.types == want.types
            for k in ("typeid", "mass", "charge", "diameter", "position", "velocity", "image", "box"):
                assert np.array_equal(getattr(got, k), getattr(want, k)), k


def test_frame_to_system_feeds_the_hot_path_layouts(tmp_path):
    """GSD (float32) -> HOOMD device layouts: the synthetic recipe stores float32-exact positions, charges and
    masses, so everything but the velocities survives the file bit for bit."""
    s = synth.make_system(5000, replica=1)
    path = str(tmp_path / "init.gsd")
    gsdio.write_gsd(path, [gsdio.system_to_frame(s)])
    with gsdio.open_gsd(path) as f:
        t = gsdio.frame_to_system(f[0])
    assert t.N == s.N and t.L_typeid == s.L_typeid and tuple(t.types) == tuple(s.types)
    n_mol = s.N - 1   # the photon's coordinate is a double draw: it is rounded to float32 by the file
    assert np.array_equal(t.pos[:n_mol].view(np.uint64), s.pos[:n_mol].view(np.uint64))      # xyz and the type bits in .w
    assert np.array_equal(t.pos[n_mol, :3], s.pos[n_mol, :3].astype(np.float32).astype(np.float64))
    assert np.array_equal(t.pos[:, 3].view(np.uint64), s.pos[:, 3].view(np.uint64))
    assert np.array_equal(t.charge, s.charge) and np.array_equal(t.image, s.image)
    assert np.array_equal(t.vel[:, 3], s.vel[:, 3])
    assert np.array_equal(t.vel[:, :3], s.vel[:, :3].astype(np.float32).astype(np.float64))
    assert np.allclose(t.box, s.box, rtol=1e-7)


def test_add_cavity_particle(tmp_path):
    """create_cavity_particle of the run script (examples/05_advanced_run.py:495-512)."""
    s = synth.make_system(10, photon="absent")
    fr = gsdio.system_to_frame(s)
    fr.types = ["O", "N"]
    out = gsdio.add_cavity_particle(fr, position=(1.0, 2.0, 3.0))
    assert out.N == 11 and out.types == ["O", "N", "L"] and fr.N == 10 and fr.types == ["O", "N"]
    assert out.typeid[-1] == 2 and out.charge[-1] == 0.0 and out.mass[-1] == 1.0 and out.diameter[-1] == 1.0
    assert np.array_equal(out.position[-1], np.float32([1, 2, 3])) and np.array_equal(out.image[-1], [0, 0, 0])
    t = gsdio.frame_to_system(out)
    assert t.L_typeid == 2 and synth.w_to_typeid(t.pos[:, 3])[-1] == 2


def test_triclinic_box_is_refused():
    fr = gsdio.Frame(N=0).fill_defaults()
    fr.box = np.float32([5, 5, 5, 0.1, 0, 0])
    with pytest.raises(ValueError):
        gsdio.frame_to_system(fr)


def test_round_trip_property(tmp_path):
    """Property test (hypothesis): any frames of the supported chunks survive write -> read bit for bit, in both
    name-list layouts; frame count, order and per-frame particle counts included."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @st.composite
    def frames(draw):
        nfr = draw(st.integers(1, 4))
        out = []
        for k in range(nfr):
            N = draw(st.integers(0, 40))
            ntypes = draw(st.integers(1, 4))
            types = [draw(st.text(alphabet="ABCLNOXYZ", min_size=1, max_size=5)) for _ in range(ntypes)]
            seed = draw(st.integers(0, 2 ** 31 - 1))
            r = np.random.default_rng(seed)
            fr = gsdio.Frame(step=draw(st.integers(0, 2 ** 40)), N=N, types=types)
            fr.box = np.float32([r.uniform(1, 50), r.uniform(1, 50), r.uniform(1, 50), 0, 0, 0])
            fr.typeid = r.integers(0, ntypes, size=N).astype(np.uint32)
            fr.mass = r.uniform(0.5, 3e4, size=N).astype(np.float32)
            fr.charge = r.normal(size=N).astype(np.float32)
            fr.diameter = np.ones(N, np.float32)
            fr.position = r.normal(scale=20, size=(N, 3)).astype(np.float32)
            fr.velocity = r.normal(scale=1e-3, size=(N, 3)).astype(np.float32)
            fr.image = r.integers(-3, 4, size=(N, 3)).astype(np.int32)
            out.append(fr)
        return out

    counter = [0]

    @settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(frames(), st.sampled_from([(1, 0), (2, 0)]))
    def check(frs, layer):
        counter[0] += 1
        path = str(tmp_path / f"p{counter[0]}.gsd")
        gsdio.write_gsd(path, frs, file_layer=layer)
        with gsdio.open_gsd(path) as f:
            assert len(f) == len(frs)
            for k, want in enumerate(frs):
                got = f[k]
                assert (got.step, got.N, got.types) == (want.step, want.N, want.types)
                for name in ("typeid", "mass", "charge", "diameter", "position", "velocity", "image", "box"):
                    a, b = getattr(got, name), getattr(want, name)
                    assert a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8)), name

    check()

"""Minimal GSD reader / writer for the chunks the hot path consumes.

The reference reads its input frame with `gsd.hoomd.open(input_gsd, 'r')[frame]` and lets HOOMD write the
trajectory the F(k,t) analysis is run on (reference examples/05_advanced_run.py:404-419, 1231-1246); the `gsd`
package is not in this image, so BASELINE configs 1 and 5 ("init-0.gsd", "a 1M-particle trajectory") could not
be consumed at all.  This module reads and writes the GSD *file layer* (header, index, name list, raw row-major
chunks) and the part of the `hoomd` schema the path needs:
    configuration/step, configuration/dimensions, configuration/box,
    particles/N, types, typeid, mass, charge, diameter, position, velocity, image.

PARITY UNPINNED: third-party format (glotzerlab/gsd, file layer 1.0 / 2.x, schema hoomd 1.4), restated from
its published specification.  Neither the gsd package nor any .gsd file exists under /root/reference or in
this image (`find / -name '*.gsd'` is empty), so the reader is checked against (a) files produced by the
writer here and (b) a file assembled byte by byte in tests/test_gsdio.py straight from the specification,
never against a file written by the real library.  Re-verify on a machine that has `gsd`.

File layer (little endian):
    header, 256 B : u64 magic 0x65DF65DF65DF65DF | u64 index_location | u64 index_allocated_entries |
                    u64 namelist_location | u64 namelist_allocated_entries | u32 schema_version |
                    u32 gsd_version | char application[64] | char schema[64] | char reserved[80]
    index entry, 32 B : u64 frame | u64 N | i64 location | u32 M | u16 id | u8 type | u8 flags
                    (location == 0: unused entry; entries are ordered by frame)
    name list     : namelist_allocated_entries * 64 B; file layer 1.0: one NUL-padded name per 64 B slot;
                    2.x: NUL-terminated names packed back to back.  id = ordinal of the name.
    chunk         : N x M values of `type`, row-major, at `location`
    type codes    : 1 u8, 2 u16, 3 u32, 4 u64, 5 i8, 6 i16, 7 i32, 8 i64, 9 f32, 10 f64
A chunk that is absent from frame i > 0 takes frame 0's value when the particle count matches, else the schema
default (mass 1, charge 0, diameter 1, typeid 0, image 0, velocity 0, types ['A'], box [1,1,1,0,0,0])."""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

MAGIC = 0x65DF65DF65DF65DF
HEADER = struct.Struct("<QQQQQII64s64s80s")
INDEX = np.dtype([("frame", "<u8"), ("N", "<u8"), ("location", "<i8"), ("M", "<u4"), ("id", "<u2"), ("type", "u1"),
                  ("flags", "u1")])
NAME_SIZE = 64
TYPES = {1: np.uint8, 2: np.uint16, 3: np.uint32, 4: np.uint64, 5: np.int8, 6: np.int16, 7: np.int32, 8: np.int64,
         9: np.float32, 10: np.float64}
CODES = {np.dtype(v): k for k, v in TYPES.items()}
assert HEADER.size == 256 and INDEX.itemsize == 32


def _version(major, minor):
    return (major << 16) | minor


@dataclass
class Frame:
    """The subset of gsd.hoomd.Frame the path uses (same attribute meaning; arrays as GSD stores them)."""
    step: int = 0
    dimensions: int = 3
    box: np.ndarray = field(default_factory=lambda: np.array([1, 1, 1, 0, 0, 0], dtype=np.float32))
    N: int = 0
    types: list = field(default_factory=lambda: ["A"])
    typeid: np.ndarray | None = None     # uint32 [N]
    mass: np.ndarray | None = None       # float32 [N]
    charge: np.ndarray | None = None     # float32 [N]
    diameter: np.ndarray | None = None   # float32 [N]
    position: np.ndarray | None = None   # float32 [N, 3]
    velocity: np.ndarray | None = None   # float32 [N, 3]
    image: np.ndarray | None = None      # int32 [N, 3]

    _PER_PARTICLE = {"typeid": (np.uint32, 1, 0), "mass": (np.float32, 1, 1.0), "charge": (np.float32, 1, 0.0),
                     "diameter": (np.float32, 1, 1.0), "position": (np.float32, 3, 0.0), "velocity": (np.float32, 3, 0.0),
                     "image": (np.int32, 3, 0)}

    def fill_defaults(self):
        for name, (dt, m, val) in self._PER_PARTICLE.items():
            if getattr(self, name) is None:
                setattr(self, name, np.full((self.N,) if m == 1 else (self.N, m), val, dtype=dt))
        return self


class GSDFile:
    """Read-only view of a GSD file: len(f), f[i] -> Frame (negative i allowed, as in
    examples/05_advanced_run.py:405-409), f.chunk(frame, name) -> raw array or None."""

    def __init__(self, path: str):
        self._fh = open(path, "rb")
        raw = self._fh.read(HEADER.size)
        if len(raw) != HEADER.size:
            raise ValueError(f"{path}: not a GSD file (shorter than the 256-byte header)")
        (magic, iloc, ialloc, nloc, nalloc, self.schema_version, self.gsd_version, app, schema, _) = HEADER.unpack(raw)
        if magic != MAGIC:
            raise ValueError(f"{path}: not a GSD file (bad magic {magic:#x})")
        if (self.gsd_version >> 16) not in (1, 2):
            raise ValueError(f"{path}: unsupported GSD file layer version {self.gsd_version >> 16}.{self.gsd_version & 0xFFFF}")
        self.application = app.split(b"\0")[0].decode()
        self.schema = schema.split(b"\0")[0].decode()
        self._fh.seek(nloc)
        block = self._fh.read(nalloc * NAME_SIZE)
        if (self.gsd_version >> 16) == 1:
            names = [block[i:i + NAME_SIZE].split(b"\0")[0] for i in range(0, len(block), NAME_SIZE)]
            names = [n for n in names if n]
        else:
            names = [n for n in block.split(b"\0") if n]
        self.names = [n.decode() for n in names]
        self._fh.seek(iloc)
        idx = np.frombuffer(self._fh.read(ialloc * INDEX.itemsize), dtype=INDEX)
        self._index = idx[idx["location"] != 0]
        self.nframes = int(self._index["frame"].max()) + 1 if len(self._index) else 0

    def close(self):
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __len__(self):
        return self.nframes

    def chunk(self, frame: int, name: str):
        if name not in self.names:
            return None
        cid = self.names.index(name)
        hit = self._index[(self._index["frame"] == frame) & (self._index["id"] == cid)]
        if len(hit) == 0:
            return None
        e = hit[0]
        if int(e["type"]) not in TYPES:
            raise ValueError(f"chunk {name}: unknown type code {int(e['type'])}")
        dt = np.dtype(TYPES[int(e["type"])]).newbyteorder("<")
        n, m = int(e["N"]), int(e["M"])
        self._fh.seek(int(e["location"]))
        raw = self._fh.read(n * m * dt.itemsize)
        if len(raw) != n * m * dt.itemsize:
            raise ValueError(f"chunk {name} of frame {frame}: truncated file")
        return np.frombuffer(raw, dtype=dt).reshape(n, m).copy()

    def _get(self, frame, name, n_expected=None):
        """frame's chunk, else frame 0's when the row count matches, else None (-> schema default)."""
        c = self.chunk(frame, name)
        if c is None and frame != 0:
            c = self.chunk(0, name)
            if c is not None and n_expected is not None and c.shape[0] != n_expected:
                c = None
        return c

    def __getitem__(self, i: int) -> Frame:
        if i < 0:
            i += self.nframes
        if not 0 <= i < self.nframes:
            raise IndexError(i)
        fr = Frame()
        c = self._get(i, "configuration/step")
        fr.step = int(c[0, 0]) if c is not None else 0
        c = self._get(i, "configuration/dimensions")
        fr.dimensions = int(c[0, 0]) if c is not None else 3
        c = self._get(i, "configuration/box")
        if c is not None:
            fr.box = c.reshape(-1).astype(np.float32)
        c = self._get(i, "particles/N")
        fr.N = int(c[0, 0]) if c is not None else 0
        c = self._get(i, "particles/types")
        if c is not None:
            fr.types = [bytes(row).split(b"\0")[0].decode() for row in c.view(np.uint8)]
        for name, (dt, m, _) in Frame._PER_PARTICLE.items():
            c = self._get(i, "particles/" + name, fr.N)
            if c is not None:
                if c.shape[0] != fr.N:
                    raise ValueError(f"particles/{name}: {c.shape[0]} rows for N = {fr.N}")
                setattr(fr, name, c.reshape(fr.N) if m == 1 else c)
        return fr.fill_defaults()


def open_gsd(path: str) -> GSDFile:
    return GSDFile(path)


def write_gsd(path: str, frames, application: str = "cavb200", file_layer=(2, 0)) -> None:
    """Write `frames` (iterable of Frame) as a hoomd-schema GSD file: header | chunks | name list | index."""
    names: list[str] = []
    entries = []
    with open(path, "wb") as fh:
        fh.write(b"\0" * HEADER.size)

        def put(frame_no, name, arr, dt, m):
            a = np.ascontiguousarray(np.asarray(arr, dtype=dt).reshape(-1, m))
            if name not in names:
                names.append(name)
            entries.append((frame_no, a.shape[0], fh.tell(), m, names.index(name), CODES[np.dtype(dt)], 0))
            fh.write(a.astype(np.dtype(dt).newbyteorder("<"), copy=False).tobytes())

        for k, fr in enumerate(frames):
            put(k, "configuration/step", [fr.step], np.uint64, 1)
            put(k, "configuration/dimensions", [fr.dimensions], np.uint8, 1)
            put(k, "configuration/box", fr.box, np.float32, 1)
            put(k, "particles/N", [fr.N], np.uint32, 1)
            width = max(len(t) for t in fr.types) + 1
            tb = np.zeros((len(fr.types), width), dtype=np.int8)
            for r, t in enumerate(fr.types):
                tb[r, :len(t)] = np.frombuffer(t.encode(), dtype=np.int8)
            put(k, "particles/types", tb, np.int8, width)
            for name, (dt, m, _) in Frame._PER_PARTICLE.items():
                v = getattr(fr, name)
                if v is not None:
                    if np.asarray(v).shape[0] != fr.N:
                        raise ValueError(f"particles/{name}: {np.asarray(v).shape[0]} rows for N = {fr.N}")
                    put(k, "particles/" + name, v, dt, m)
        nloc = fh.tell()
        if file_layer[0] == 1:
            block = b"".join(n.encode().ljust(NAME_SIZE, b"\0") for n in names)
            nalloc = len(names)
        else:
            packed = b"".join(n.encode() + b"\0" for n in names)
            nalloc = (len(packed) + NAME_SIZE) // NAME_SIZE  # at least one trailing NUL
            block = packed.ljust(nalloc * NAME_SIZE, b"\0")
        fh.write(block)
        iloc = fh.tell()
        idx = np.zeros(len(entries) + 1, dtype=INDEX)  # one unused (location 0) entry terminates the list
        entries.sort(key=lambda e: (e[0], e[4]))
        for r, e in enumerate(entries):
            idx[r] = e
        fh.write(idx.tobytes())
        fh.seek(0)
        fh.write(HEADER.pack(MAGIC, iloc, len(idx), nloc, nalloc, _version(1, 4), _version(*file_layer),
                             application.encode().ljust(64, b"\0"), b"hoomd".ljust(64, b"\0"), b"\0" * 80))


# ---- bridge to the hot path's array layouts -------------------------------------------------------
def add_cavity_particle(fr: Frame, position=(0.0, 0.0, 0.0)) -> Frame:
    """create_cavity_particle (reference examples/05_advanced_run.py:495-512): append type 'L' (typeid of 'L'),
    charge 0, mass 1, diameter 1, image 0."""
    out = Frame(**{k: (v.copy() if isinstance(v, np.ndarray) else (list(v) if isinstance(v, list) else v))
                   for k, v in fr.__dict__.items()})
    if "L" not in out.types:
        out.types.append("L")
    out.N += 1
    out.typeid = np.append(out.typeid, np.uint32(out.types.index("L")))
    out.position = np.append(out.position, np.asarray([position], dtype=np.float32), axis=0)
    out.velocity = np.append(out.velocity, np.zeros((1, 3), dtype=np.float32), axis=0)
    out.charge = np.append(out.charge, np.float32(0.0))
    out.mass = np.append(out.mass, np.float32(1.0))
    out.diameter = np.append(out.diameter, np.float32(1.0))
    out.image = np.vstack([out.image, np.zeros((1, 3), dtype=np.int32)])
    return out


def frame_to_system(fr: Frame):
    """GSD frame -> synth.System in HOOMD's device layouts: pos double4 {x,y,z,type bits}, vel double4
    {vx,vy,vz,mass}, charge double, image int3; float32 values widened exactly."""
    from . import synth
    if abs(float(fr.box[3])) + abs(float(fr.box[4])) + abs(float(fr.box[5])) != 0.0:
        raise ValueError("triclinic box: the cavity force is defined for orthorhombic boxes only "
                         "(reference src/CavityForceCompute.cc:97 uses box.getL())")
    pos = np.zeros((fr.N, 4), dtype=np.float64)
    pos[:, :3] = fr.position.astype(np.float64)
    pos[:, 3] = synth.typeid_to_w(fr.typeid)
    vel = np.zeros((fr.N, 4), dtype=np.float64)
    vel[:, :3] = fr.velocity.astype(np.float64)
    vel[:, 3] = fr.mass.astype(np.float64)
    L_typeid = fr.types.index("L") if "L" in fr.types else -1
    return synth.System(pos, vel, fr.charge.astype(np.float64), np.ascontiguousarray(fr.image, dtype=np.int32),
                        tuple(float(x) for x in fr.box[:3]), L_typeid, tuple(fr.types))


def system_to_frame(s, step: int = 0) -> Frame:
    """synth.System -> GSD frame (float32 storage, as HOOMD's GSD writer does)."""
    from . import synth
    fr = Frame(step=step, N=s.N, types=list(s.types))
    fr.box = np.array([s.box[0], s.box[1], s.box[2], 0, 0, 0], dtype=np.float32)
    fr.position = s.pos[:, :3].astype(np.float32)
    fr.velocity = s.vel[:, :3].astype(np.float32)
    fr.mass = s.vel[:, 3].astype(np.float32)
    fr.charge = s.charge.astype(np.float32)
    fr.typeid = synth.w_to_typeid(s.pos[:, 3]).astype(np.uint32)
    fr.image = np.ascontiguousarray(s.image, dtype=np.int32)
    return fr.fill_defaults()

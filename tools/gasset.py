"""Reader/writer for the engine's asset container, `.gmap` maps and `.gmdl` models.

Test/fixture tooling only (the product's loader is the C one in
`c-game-engine_b200/host/gpx_assets.c`).  Format follows the reference loaders:

  container : engine/src/assets/AssetReader.c:150-257, AssetReader.h:15-17
              <I magic 0x454D4147><B ver=2><B type><B typeVer><Q rawSize><Q gzSize> + gzip body
  scalars   : engine/src/assets/DataReader.c:40-100 (little endian, size_t = 8 bytes,
              strings = <Q len> + len bytes)
  params    : engine/src/structs/KVList.c:19-76,237-253, KVList.h:40-54
  map       : engine/src/assets/MapLoader.c:40-314
  model     : engine/src/assets/ModelLoader.c:33-211
"""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass, field

import numpy as np

MAGIC = 0x454D4147
VERSION = 2
HEADER = struct.Struct("<IBBBQQ")  # 23 bytes


class Reader:
    def __init__(self, data: bytes):
        self.d = data
        self.o = 0

    def take(self, fmt: str):
        s = struct.Struct("<" + fmt)
        if self.o + s.size > len(self.d):
            raise ValueError("DataReader Buffer Overrun")
        v = s.unpack_from(self.d, self.o)
        self.o += s.size
        return v if len(v) > 1 else v[0]

    def raw(self, n: int) -> bytes:
        if self.o + n > len(self.d):
            raise ValueError("DataReader Buffer Overrun")
        b = self.d[self.o:self.o + n]
        self.o += n
        return b

    def string(self) -> str:
        n = self.take("Q")
        return self.raw(n).split(b"\0", 1)[0].decode("utf-8", "replace")

    @property
    def left(self) -> int:
        return len(self.d) - self.o


def read_container(blob: bytes):
    """-> (type, typeVersion, decompressed bytes)."""
    magic, ver, typ, tver, raw_size, gz_size = HEADER.unpack_from(blob, 0)
    if magic != MAGIC:
        raise ValueError("bad magic")
    if ver != VERSION:
        raise ValueError("bad container version")
    if len(blob) - HEADER.size != gz_size:
        raise ValueError("compressedSize mismatch")
    body = zlib.decompress(blob[HEADER.size:], zlib.MAX_WBITS | 16)
    if len(body) != raw_size:
        raise ValueError("decompressedSize mismatch")
    return typ, tver, body


def write_container(typ: int, tver: int, body: bytes) -> bytes:
    co = zlib.compressobj(9, zlib.DEFLATED, zlib.MAX_WBITS | 16)
    gz = co.compress(body) + co.flush()
    return HEADER.pack(MAGIC, VERSION, typ, tver, len(body), len(gz)) + gz


# ---------------------------------------------------------------- params
P_BYTE, P_INT, P_FLOAT, P_BOOL, P_STRING, P_NONE, P_COLOR, P_KVLIST, P_ARRAY, P_U64, P_VEC2, P_VEC3 = range(12)


def read_param(r: Reader):
    t = r.take("B")
    if t == P_BYTE:
        return t, r.take("B")
    if t == P_INT:
        return t, r.take("i")
    if t == P_FLOAT:
        return t, r.take("f")
    if t == P_BOOL:
        return t, r.take("B") != 0
    if t == P_COLOR:
        return t, r.take("4f")
    if t == P_STRING:
        return t, r.string()
    if t == P_ARRAY:
        n = r.take("Q")
        return t, [read_param(r) for _ in range(n)]
    if t == P_KVLIST:
        return t, read_kvlist(r)
    if t == P_U64:
        return t, r.take("Q")
    if t == P_VEC2:
        return t, r.take("2f")
    if t == P_VEC3:
        return t, r.take("3f")
    return t, None


def read_kvlist(r: Reader) -> dict:
    n = r.take("Q")
    out = {}
    for _ in range(n):
        k = r.string()
        out[k] = read_param(r)
    return out


def _wstr(s: str) -> bytes:
    b = s.encode() + b"\0"
    return struct.pack("<Q", len(b)) + b


def write_param(p) -> bytes:
    t, v = p
    b = struct.pack("<B", t)
    if t == P_BYTE:
        b += struct.pack("<B", v)
    elif t == P_INT:
        b += struct.pack("<i", v)
    elif t == P_FLOAT:
        b += struct.pack("<f", v)
    elif t == P_BOOL:
        b += struct.pack("<B", 1 if v else 0)
    elif t == P_COLOR:
        b += struct.pack("<4f", *v)
    elif t == P_STRING:
        b += _wstr(v)
    elif t == P_ARRAY:
        b += struct.pack("<Q", len(v)) + b"".join(write_param(x) for x in v)
    elif t == P_KVLIST:
        b += write_kvlist(v)
    elif t == P_U64:
        b += struct.pack("<Q", v)
    elif t == P_VEC2:
        b += struct.pack("<2f", *v)
    elif t == P_VEC3:
        b += struct.pack("<3f", *v)
    return b


def write_kvlist(kv: dict) -> bytes:
    b = struct.pack("<Q", len(kv))
    for k, p in kv.items():
        b += _wstr(k) + write_param(p)
    return b


# ---------------------------------------------------------------- map
@dataclass
class MapActor:
    cls: str
    pos: tuple
    euler: tuple
    connections: list
    params: dict


@dataclass
class CollisionMesh:
    pos: np.ndarray                       # (3,) f32, mesh origin
    subshapes: list                       # list of (T,3,3) f32 arrays, vertices relative to pos


@dataclass
class GMap:
    sky: str | None = None
    rpc_icon: str = ""
    rpc_name: str = ""
    actors: list = field(default_factory=list)
    models: list = field(default_factory=list)      # (material, verts(n,7) f32, indices u32)
    meshes: list = field(default_factory=list)      # CollisionMesh
    lightmap: tuple = (0, 0, b"")
    lights: bytes = b""
    n_lights: int = 0
    leftover: int = 0

    def tri_counts(self):
        return [int(sum(len(s) for s in m.subshapes)) for m in self.meshes]


def parse_gmap(body: bytes) -> GMap:
    r = Reader(body)
    m = GMap()
    if r.take("B"):
        m.sky = r.string()
    m.rpc_icon = r.string()
    m.rpc_name = r.string()
    for _ in range(r.take("Q")):
        cls = r.string()
        pos = r.take("3f")
        eul = r.take("3f")
        conns = []
        for _ in range(r.take("Q")):
            out_name, tgt, tgt_in = r.string(), r.string(), r.string()
            override = read_param(r) if r.take("B") else None
            refires = r.take("Q")
            conns.append((out_name, tgt, tgt_in, override, refires))
        params = read_kvlist(r)
        m.actors.append(MapActor(cls, pos, eul, conns, params))
    for _ in range(r.take("Q")):
        mat = r.string()
        nv = r.take("I")
        verts = np.frombuffer(r.raw(nv * 28), dtype="<f4").reshape(nv, 7).copy()
        ni = r.take("I")
        idx = np.frombuffer(r.raw(ni * 4), dtype="<u4").copy()
        m.models.append((mat, verts, idx))
    for _ in range(r.take("Q")):
        pos = np.array(r.take("3f"), dtype=np.float32)
        subs = []
        for _ in range(r.take("Q")):
            nt = r.take("Q")
            subs.append(np.frombuffer(r.raw(nt * 36), dtype="<f4").reshape(nt, 3, 3).copy())
        m.meshes.append(CollisionMesh(pos, subs))
    w, h = r.take("QQ")
    m.lightmap = (w, h, r.raw(2 * 4 * w * h))
    if r.left >= 2:
        m.n_lights = r.take("H")
        m.lights = r.raw(min(r.left, m.n_lights * 36))
    m.leftover = r.left
    return m


def load_gmap(path: str) -> GMap:
    with open(path, "rb") as f:
        _, _, body = read_container(f.read())
    return parse_gmap(body)


def build_gmap_body(meshes, actors=(), lightmap=(0, 0, b"")) -> bytes:
    """Serialise a map that carries only what the physics path reads (no render models, no lights)."""
    b = struct.pack("<B", 0) + _wstr("icon") + _wstr("synthetic")
    b += struct.pack("<Q", len(actors))
    for a in actors:
        b += _wstr(a.cls) + struct.pack("<6f", *a.pos, *a.euler)
        b += struct.pack("<Q", len(a.connections))
        for (o, t, ti, ov, rf) in a.connections:
            b += _wstr(o) + _wstr(t) + _wstr(ti)
            b += struct.pack("<B", 1) + write_param(ov) if ov is not None else struct.pack("<B", 0)
            b += struct.pack("<Q", rf)
        b += write_kvlist(a.params)
    b += struct.pack("<Q", 0)  # render models
    b += struct.pack("<Q", len(meshes))
    for cm in meshes:
        b += struct.pack("<3f", *[float(x) for x in cm.pos]) + struct.pack("<Q", len(cm.subshapes))
        for s in cm.subshapes:
            s = np.ascontiguousarray(s, dtype="<f4")
            b += struct.pack("<Q", len(s)) + s.tobytes()
    w, h, px = lightmap
    b += struct.pack("<QQ", w, h) + px
    b += struct.pack("<H", 0)
    return b


MAP_ASSET_TYPE = 0  # only the container checks magic/version; the type byte is carried through


# ---------------------------------------------------------------- model
@dataclass
class GModel:
    collision_type: int = 0                          # 0 none, 1 static mesh, 2 dynamic hulls
    bb_origin: tuple = (0, 0, 0)
    bb_extents: tuple = (0, 0, 0)
    hulls: list = field(default_factory=list)        # (offset(3,), points(n,3))
    tris: np.ndarray | None = None                   # (T,3,3)
    leftover: int = 0


def parse_gmdl(body: bytes) -> GModel:
    r = Reader(body)
    g = GModel()
    n_mat, n_slot, n_skin, n_lod = r.take("4I")
    g.collision_type = r.take("B")
    for _ in range(n_mat):
        r.string()
        r.raw(16 + 4)
    r.raw(4 * n_slot * n_skin)
    for _ in range(n_lod):
        r.raw(8)
        nv = r.take("Q")
        r.raw(nv * 48)
        r.take("I")
        counts = r.take(f"{n_slot}I") if n_slot != 1 else (r.take("I"),)
        for c in counts:
            r.raw(4 * c)
    g.bb_origin = r.take("3f")
    g.bb_extents = r.take("3f")
    if g.collision_type == 2:
        for _ in range(r.take("Q")):
            n = r.take("Q")
            off = np.array(r.take("3f"), dtype=np.float32)
            pts = np.frombuffer(r.raw(n * 12), dtype="<f4").reshape(n, 3).copy()
            g.hulls.append((off, pts))
    elif g.collision_type == 1:
        n = r.take("Q")
        g.tris = np.frombuffer(r.raw(n * 36), dtype="<f4").reshape(n, 3, 3).copy()
    g.leftover = r.left
    return g


def load_gmdl(path: str) -> GModel:
    with open(path, "rb") as f:
        _, _, body = read_container(f.read())
    return parse_gmdl(body)


def build_gmdl_body(collision_type, bb_origin, bb_extents, hulls=(), tris=None) -> bytes:
    """Serialise a model with no render data (0 materials/skins/lods) + its collision section."""
    b = struct.pack("<4IB", 0, 0, 0, 0, collision_type)
    b += struct.pack("<6f", *bb_origin, *bb_extents)
    if collision_type == 2:
        b += struct.pack("<Q", len(hulls))
        for off, pts in hulls:
            pts = np.ascontiguousarray(pts, dtype="<f4")
            b += struct.pack("<Q", len(pts)) + struct.pack("<3f", *[float(x) for x in off]) + pts.tobytes()
    elif collision_type == 1:
        t = np.ascontiguousarray(tris, dtype="<f4")
        b += struct.pack("<Q", len(t)) + t.tobytes()
    return b

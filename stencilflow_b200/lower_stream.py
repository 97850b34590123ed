"""Streamed (temporally blocked) lowering: several chained operators in one pass over HBM.

What the reference builds out of FPGA shift registers and FIFOs -- every operator a processing
element holding a sliding window of its inputs (``stencilflow/stencil/intel_fpga.py:58-69,367-461``),
connected by delay buffers sized by ``compute_delay_buffer`` (``stencilflow/kernel_chain_graph.py:476-559``)
so that only program inputs and outputs touch off-chip memory (``generate_sdfg``,
``stencilflow/sdfg_generator.py:219-577``) -- becomes one CUDA kernel per fusion group:

* the group streams along the outermost dimension (i for 3-D programs, j for 2-D ones); a CTA owns a
  tile of the remaining dimension(s) including the halo the fused operators consume;
* planes of the group's input fields are staged into a shared-memory ring by TMA
  (``cp.async.bulk.tensor`` + mbarrier), ``PREFETCH`` steps ahead of their use;
* every field (group input or operator result) lives in a *register sliding window* over the streamed
  dimension: thread (warp, lane) owns R rows x V consecutive cells and keeps its own values of the
  last W planes, so taps at (dj, dk) = (0, 0) cost nothing;
* in-plane neighbours along the innermost dimension come from warp shuffles, neighbours along the
  row dimension from the adjacent warps through small shared-memory exchange rings (only the edge
  rows of each warp are published);
* operator A runs ``lag(A)`` planes behind the input stream -- the plane lags are the restriction of
  the reference's path-length recurrence to the streamed dimension with unit latency -- and one
  ``__syncthreads`` per streamed plane separates producers from consumers;
* boundary conditions are applied where a field is *produced*: cells outside the domain are set to
  the constant (or to -100000 for ``shrink``) the consumers would read, which is exactly what
  ``ExpandStencilCPU`` selects per tap (``stencilflow/stencil/cpu.py:73-102``).
"""

import collections
import hashlib
import math
from typing import Dict, List, Optional, Tuple

from . import dtypes
from . import expr as ex
from .lower_cuda import (KernelSpec, LaunchSpec, LoweredProgram, _MATH_F32, _MATH_F64, ctype_of,
                         literal)
from .stencil_op import JUNK_VAL, StencilOp, StencilProgram

SMEM_LIMIT = 227 * 1024
REG_BUDGET = 110          # estimated window registers per thread the planner accepts

STREAM_PRELUDE = r"""
// ---- streamed-kernel support (TMA + mbarrier, sm_100a) ----
struct __align__(64) CUtensorMap_st { unsigned long long opaque[16]; };
typedef CUtensorMap_st CUtensorMap;

__device__ __forceinline__ u32 sf_smem_addr(const void* p) {
    return (u32)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void sf_mbar_init(void* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sf_smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void sf_fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void sf_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void sf_mbar_expect_tx(void* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sf_smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void sf_mbar_wait(void* bar, u32 parity) {
    u32 addr = sf_smem_addr(bar);
    u32 done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void sf_tma_load_3d(void* dst, const CUtensorMap* map, void* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(sf_smem_addr(dst)), "l"(map), "r"(sf_smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void sf_tma_load_2d(void* dst, const CUtensorMap* map, void* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(sf_smem_addr(dst)), "l"(map), "r"(sf_smem_addr(bar)), "r"(c0), "r"(c1)
        : "memory");
}
"""


class NotStreamable(Exception):
    pass


class _FieldInfo:
    def __init__(self, name, kind, dtype):
        self.name = name
        self.kind = kind              # "ext" (read from HBM) | "op" (produced in the group)
        self.dtype = dtype
        self.lag = 0                  # planes behind the stream head when produced
        self.window = 1               # planes kept in registers
        self.row_reach = 0            # max |dj| over consumers  -> rows published per side
        self.col_reach = 0            # max |dk| over consumers
        self.row_ring = 0             # exchange ring depth (0 = not published)
        self.col_ring = 0
        self.bc = None                # value consumers read outside the domain (None = never read there)
        self.stored = False
        self.consumed = False
        self.back = 0                 # planes of history a chunk needs before its first output plane
        self.fwd = 0                  # planes beyond the last output plane of a chunk that are read
        self.need = [0, 0, 0, 0]      # halo (row lo, row hi, col lo, col hi) this field must be valid on


class GroupAnalysis:
    """Lags, windows, rings and halos of a fusion group (see the module docstring)."""

    def __init__(self, program: StencilProgram, ops: List[StencilOp], exchange_cols: bool):
        self.program = program
        self.ops = ops
        self.ndim = len(program.shape)
        if self.ndim not in (2, 3):
            raise NotStreamable("only 2-D and 3-D programs stream")
        self.exchange_cols = exchange_cols
        self.dtype = ops[0].data_type
        if self.dtype not in (dtypes.float32, dtypes.float64):
            raise NotStreamable("only float32/float64 fields stream")
        self.fields: Dict[str, _FieldInfo] = collections.OrderedDict()
        self.taps: Dict[str, List[Tuple[str, int, int, int]]] = {}
        produced = {op.name for op in ops}
        later = set()
        seen_group = False
        for op in program.ops:
            if op in ops:
                seen_group = True
                continue
            if seen_group or True:
                later.update(f for f in op.accesses if f in produced)
        for op in ops:
            if op.data_type != self.dtype:
                raise NotStreamable("mixed result types in group")
            for s in op.scalars:
                pass
            taps = []
            for field in op.accesses:
                f = program.fields[field]
                if f.data_type != self.dtype:
                    raise NotStreamable("field {} has a different type".format(field))
                if list(f.dims) != list(program.iterators):
                    raise NotStreamable("lower-dimensional array input {}".format(field))
                if field not in self.fields:
                    if field in produced:
                        raise NotStreamable("operators out of order")
                    self.fields[field] = _FieldInfo(field, "ext", f.data_type)
                for off in op.offsets3(field):
                    if self.ndim == 3:
                        d, dj, dk = off
                    else:
                        d, dj, dk = off[1], 0, off[2]
                    taps.append((field, d, dj, dk))
                bc = op.boundary_conditions.get(field)
                if any((t[1], t[2], t[3]) != (0, 0, 0) for t in taps if t[0] == field):
                    if bc is None:
                        raise NotStreamable("missing boundary condition")
                    if bc["btype"] == "copy":
                        raise NotStreamable("copy boundary")
                    val = float(bc["value"]) if bc["btype"] == "constant" else float(JUNK_VAL)
                    info = self.fields[field]
                    if info.bc is not None and info.bc != val:
                        raise NotStreamable("consumers of {} disagree on the boundary value".format(field))
                    info.bc = val
            self.taps[op.name] = taps
            info = _FieldInfo(op.name, "op", op.data_type)
            info.stored = (program.fields[op.name].kind == "output") or (op.name in later)
            self.fields[op.name] = info
        for op in ops:
            for (field, d, dj, dk) in self.taps[op.name]:
                self.fields[field].consumed = True
        for op in ops:
            info = self.fields[op.name]
            if not info.stored and not info.consumed:
                raise NotStreamable("dead operator {}".format(op.name))
        self._lags()
        self._halos()

    def _is_exchange(self, dj, dk):
        return dj != 0 or (self.exchange_cols and dk != 0)

    def _lags(self):
        f = self.fields
        for op in self.ops:
            lag = 0
            for (field, d, dj, dk) in self.taps[op.name]:
                lag = max(lag, f[field].lag + d + (1 if self._is_exchange(dj, dk) else 0), f[field].lag)
            f[op.name].lag = lag
        # delay the loading of inputs that are only needed late (keeps their windows short)
        for name, info in f.items():
            if info.kind != "ext":
                continue
            slack = None
            for op in self.ops:
                for (field, d, dj, dk) in self.taps[op.name]:
                    if field == name:
                        s = f[op.name].lag - d - (1 if self._is_exchange(dj, dk) else 0)
                        slack = s if slack is None else min(slack, s)
            info.lag = max(0, slack or 0)
        for op in self.ops:
            for (field, d, dj, dk) in self.taps[op.name]:
                src = f[field]
                age = f[op.name].lag - d - src.lag
                assert age >= 0
                src.window = max(src.window, age + 1)
                src.row_reach = max(src.row_reach, abs(dj))
                src.col_reach = max(src.col_reach, abs(dk))
                if dj != 0:
                    assert age >= 1
                    src.row_ring = max(src.row_ring, age + 1)
                if self.exchange_cols and dk != 0:
                    assert age >= 1
                    src.col_ring = max(src.col_ring, age + 1)
        # history needed before the first output plane of a chunk
        for op in reversed(self.ops):
            for (field, d, dj, dk) in self.taps[op.name]:
                f[field].back = max(f[field].back, f[op.name].back + max(0, -d))
                f[field].fwd = max(f[field].fwd, f[op.name].fwd + max(0, d))

    def _halos(self):
        f = self.fields
        for op in reversed(self.ops):
            need = f[op.name].need
            for (field, d, dj, dk) in self.taps[op.name]:
                n = f[field].need
                n[0] = max(n[0], need[0] + max(0, -dj))
                n[1] = max(n[1], need[1] + max(0, dj))
                n[2] = max(n[2], need[2] + max(0, -dk))
                n[3] = max(n[3], need[3] + max(0, dk))
        ext = [i for i in f.values() if i.kind == "ext"]
        self.halo = [max(i.need[s] for i in ext) for s in range(4)]

    @property
    def ext_fields(self):
        return [i for i in self.fields.values() if i.kind == "ext"]

    @property
    def max_lag(self):
        return max(i.lag for i in self.fields.values())

    def t_begin_offset(self):
        """first step = chunk_begin + this (<= 0)"""
        return min(-i.back + i.lag for i in self.fields.values())

    def t_end_offset(self):
        """last step + 1 = chunk_end + this"""
        return max(i.lag for i in self.fields.values() if i.stored)

    def window_registers(self, R, V):
        per = self.dtype.bytes // 4
        return sum(i.window for i in self.fields.values() if i.consumed) * R * V * per


class Geometry:
    def __init__(self, ana: GroupAnalysis, V, R, WR, WC, prefetch, KS=32):
        """``KS`` = threads per tile row.  32 (a warp per row, k-neighbours by shuffle) in general;
        for groups without k-taps on a narrow innermost dimension the whole extent is one row of
        ``KS = NK / V`` threads and thread t owns row group t / KS ("flat lanes": no idle lanes)."""
        self.V, self.R, self.WR, self.WC, self.KS = V, R, WR, WC, KS
        self.NT = KS * WR * WC
        if KS != 32 and (WC != 1 or self.NT % 32 or any(i.col_reach for i in ana.fields.values())):
            raise NotStreamable("flat lanes need a single column tile and no k-taps")
        if self.NT > 1024:
            raise NotStreamable("too many threads")
        self.TR, self.TC = WR * R, WC * KS * V
        h = ana.halo
        self.HJ0, self.HJ1 = h[0], h[1]
        self.HK0 = -(-h[2] // V) * V
        self.HK1 = -(-h[3] // V) * V
        self.BJ = self.TR - self.HJ0 - self.HJ1
        self.BK = self.TC - self.HK0 - self.HK1
        self.P = prefetch
        self.D = prefetch + 1
        if ana.ndim == 2:
            self.BJ = 1
        if self.BJ < 1 or self.BK < V:
            raise NotStreamable("tile smaller than its halo")
        for i in ana.fields.values():
            if i.row_reach > R:
                raise NotStreamable("row reach exceeds rows per thread")
            if i.col_reach > V:
                raise NotStreamable("column reach exceeds the vector width")
        self.box_cols = min(self.TC, 256)
        if self.TC % self.box_cols:
            raise NotStreamable("tile width not a multiple of the TMA box")
        self.smem = self._smem(ana)

    def _smem(self, ana):
        b = ana.dtype.bytes
        off = 0
        self.tile_off, self.xrow_off, self.xcol_off = {}, {}, {}
        for i in ana.ext_fields:
            self.tile_off[i.name] = off
            off += self.D * self.TR * self.TC * b
            off = (off + 127) & ~127
        for i in ana.fields.values():
            if i.row_ring:
                self.xrow_off[i.name] = off
                off += i.row_ring * self.WR * 2 * i.row_reach * self.TC * b
                off = (off + 127) & ~127
            if i.col_ring:
                self.xcol_off[i.name] = off
                off += i.col_ring * self.WR * self.WC * 2 * self.R * self.V * b
                off = (off + 127) & ~127
        self.bar_off = off
        off += 8 * self.D
        return off


def _fmt_off(x):
    return str(x).replace("-", "m")


class StreamKernelGen:
    def __init__(self, program: StencilProgram, ops: List[StencilOp], ana: GroupAnalysis, geo: Geometry,
                 specialize=None):
        self.program, self.ops, self.ana, self.geo = program, ops, ana, geo
        self.specialize = specialize or {}
        self.ct = ana.dtype
        self.T = ctype_of(self.ct)
        self.NI, self.NJ, self.NK = program.shape3
        self.NS = program.shape[0]                     # extent of the streamed dimension
        self.scalars = []
        for op in ops:
            for s in op.scalars:
                if s not in program.constants and s not in self.specialize and s not in self.scalars:
                    self.scalars.append(s)
        self.lines: List[str] = []

    # ------------------------------------------------------------------ small emit helpers
    def emit(self, text, indent=1):
        self.lines.append("  " * indent + text)

    def lit(self, v):
        return literal(v, self.ct)

    def _slot_expr(self, ring_var, ring, age):
        """index of the ring slot written ``age`` steps ago (ring_var = slot written this step)."""
        if age == 0:
            return ring_var
        return "(({rv} + {k}) % {n})".format(rv=ring_var, k=ring - (age % ring), n=ring)

    # ------------------------------------------------------------------ kernel text
    def generate(self):
        g, a = self.geo, self.ana
        T, V, R = self.T, g.V, g.R
        e = self.emit
        ndim = a.ndim
        ext = a.ext_fields
        stored = [i for i in a.fields.values() if i.stored]
        params = ["const __grid_constant__ CUtensorMap tm_{}".format(n) for n in range(len(ext))]
        params += ["{}* __restrict__ o_{}".format(T, n) for n in range(len(stored))]
        self.sc_name = {s: "s{}".format(n) for n, s in enumerate(self.scalars)}
        params += ["const {} {}".format(ctype_of(self.program.fields[s].data_type), self.sc_name[s])
                   for s in self.scalars]
        params += ["const int s_base", "const int s_begin", "const int s_end", "const int chunk"]
        self.fid = {name: "f{}".format(n) for n, name in enumerate(a.fields)}

        e("extern __shared__ __align__(1024) unsigned char sf_smem[];")
        if g.KS == 32:
            e("const int lane = threadIdx.x & 31;")
            e("const int warp = threadIdx.x >> 5;")
        else:
            e("const int lane = threadIdx.x % {};          // column slot within the row".format(g.KS))
            e("const int warp = threadIdx.x / {};          // row group".format(g.KS))
        e("const int wr = warp / {};".format(g.WC))
        e("const int wc = warp % {};".format(g.WC))
        e("(void)wr; (void)wc;")
        e("const int tile_k0 = blockIdx.x * {};".format(g.BK))
        if ndim == 3:
            e("const int tile_j0 = blockIdx.y * {};".format(g.BJ))
        e("const int c_begin = s_begin + blockIdx.z * chunk;")
        e("const int c_end = min(c_begin + chunk, s_end);")
        e("if (c_begin >= c_end) return;")
        e("const int c0 = (wc * {} + lane) * {};            // first owned column inside the tile".format(g.KS, V))
        e("const int gk = tile_k0 - {} + c0;                 // its global k".format(g.HK0))
        if ndim == 3:
            e("const int r0 = wr * {};".format(R))
            e("const int gj0 = tile_j0 - {} + r0;".format(g.HJ0))
        # in-domain mask of the owned cells, store mask of the owned rows
        e("u32 cmask = 0;")
        e("u32 smask = 0;")
        e("#pragma unroll")
        e("for (int r = 0; r < {}; ++r) {{".format(R))
        if ndim == 3:
            e("const bool rin = (gj0 + r) >= 0 && (gj0 + r) < {};".format(self.NJ), 2)
            e("const bool rst = rin && (r0 + r) >= {} && (r0 + r) < {};".format(g.HJ0, g.TR - g.HJ1), 2)
        else:
            e("const bool rin = true, rst = true;", 2)
        e("#pragma unroll", 2)
        e("for (int v = 0; v < {}; ++v)".format(V), 2)
        e("if (rin && (gk + v) >= 0 && (gk + v) < {}) cmask |= 1u << (r * {} + v);".format(self.NK, V), 3)
        e("if (rst && c0 >= {} && c0 < {} && gk < {}) smask |= 1u << r;".format(g.HK0, g.TC - g.HK1, self.NK), 2)
        e("}")
        full_mask = (1 << (R * V)) - 1
        e("const bool interior = __syncthreads_and(cmask == {}u);".format(full_mask))
        # shared memory carve-up
        for n, i in enumerate(ext):
            e("{T}* const tile_{f} = reinterpret_cast<{T}*>(sf_smem + {o});".format(
                T=T, f=self.fid[i.name], o=g.tile_off[i.name]))
        for i in a.fields.values():
            if i.row_ring:
                e("{T}* const xrow_{f} = reinterpret_cast<{T}*>(sf_smem + {o});".format(
                    T=T, f=self.fid[i.name], o=g.xrow_off[i.name]))
            if i.col_ring:
                e("{T}* const xcol_{f} = reinterpret_cast<{T}*>(sf_smem + {o});".format(
                    T=T, f=self.fid[i.name], o=g.xcol_off[i.name]))
        e("unsigned long long* const bars = reinterpret_cast<unsigned long long*>(sf_smem + {});".format(g.bar_off))
        e("if (threadIdx.x == 0) {")
        e("for (int s = 0; s < {}; ++s) sf_mbar_init(&bars[s], 1);".format(g.D), 2)
        e("sf_fence_barrier_init();", 2)
        e("}")
        e("__syncthreads();")
        # register windows
        for i in a.fields.values():
            if i.consumed:
                e("{T} w_{f}[{W}][{R}][{V}];".format(T=T, f=self.fid[i.name], W=i.window, R=R, V=V))
                e("#pragma unroll")
                e("for (int a = 0; a < {}; ++a)".format(i.window))
                e("#pragma unroll", 2)
                e("for (int r = 0; r < {}; ++r)".format(R), 2)
                e("#pragma unroll", 3)
                e("for (int v = 0; v < {}; ++v) w_{}[a][r][v] = {};".format(V, self.fid[i.name], self.lit(0)), 3)
            if i.row_ring:
                e("int xr_{} = 0;".format(self.fid[i.name]))
            if i.col_ring:
                e("int xc_{} = 0;".format(self.fid[i.name]))
        tile_bytes = g.TR * g.TC * self.ct.bytes
        e("const int t_begin = c_begin + ({});".format(a.t_begin_offset()))
        e("const int t_end = c_end + ({});".format(a.t_end_offset()))
        e("int slot = 0; u32 phase = 0;")
        # TMA issue helper as a lambda
        e("auto issue = [&](int t, int s) {")
        e("sf_mbar_expect_tx(&bars[s], {});".format(tile_bytes * len(ext)), 2)
        nbox = g.TC // g.box_cols
        for n, i in enumerate(ext):
            plane = "t - ({}) - s_base".format(i.lag)
            for b in range(nbox):
                dst = "tile_{f} + s * {sz} + {bo}".format(f=self.fid[i.name], sz=g.TR * g.TC, bo=b * g.box_cols)
                if ndim == 3:
                    e("sf_tma_load_3d({}, &tm_{}, &bars[s], tile_k0 - {} + {}, tile_j0 - {}, {});".format(
                        dst, n, g.HK0, b * g.box_cols, g.HJ0, plane), 2)
                else:
                    e("sf_tma_load_2d({}, &tm_{}, &bars[s], tile_k0 - {} + {}, {});".format(
                        dst, n, g.HK0, b * g.box_cols, plane), 2)
        e("};")
        e("if (threadIdx.x == 0) {")
        e("for (int p = 0; p < {}; ++p) if (t_begin + p < t_end) issue(t_begin + p, p);".format(g.P), 2)
        e("}")
        e("for (int t = t_begin; t < t_end; ++t) {")
        e("if (threadIdx.x == 0 && t + {P} < t_end) issue(t + {P}, (slot + {P}) % {D});".format(P=g.P, D=g.D), 2)
        e("sf_mbar_wait(&bars[slot], phase);", 2)
        for i in ext:
            self._produce_ext(i)
        for op in self.ops:
            self._produce_op(op)
        e("__syncthreads();", 2)
        e("if (++slot == {}) {{ slot = 0; phase ^= 1; }}".format(g.D), 2)
        for i in a.fields.values():
            if i.row_ring:
                e("if (++xr_{f} == {n}) xr_{f} = 0;".format(f=self.fid[i.name], n=i.row_ring), 2)
            if i.col_ring:
                e("if (++xc_{f} == {n}) xc_{f} = 0;".format(f=self.fid[i.name], n=i.col_ring), 2)
        e("}")
        body = "\n".join(self.lines)
        digest = hashlib.sha1((body + ";".join(params)).encode()).hexdigest()[:12]
        name = "sf_stream_{}".format(digest)
        src = ("extern \"C\" __global__ void __launch_bounds__({}, 1)\n{}({})\n{{\n{}\n}}\n".format(
            g.NT, name, ", ".join(params), body))
        args = [("tmap", {"field": i.name, "box": self._box()}) for i in ext]
        args += [("buf", i.name) for i in stored]
        args += [("scalar", self.program.fields[s].data_type, s) for s in self.scalars]
        args += [("slab",), ("chunk",)]
        return name, src, args

    def _box(self):
        g = self.geo
        if self.ana.ndim == 3:
            return [g.box_cols, g.TR, 1]
        return [g.box_cols, 1]

    # ------------------------------------------------------------------ producing a field
    def _finish_field(self, info: _FieldInfo, plane_expr: str):
        """``nv`` holds the new plane: apply the boundary value, rotate the window, publish edges."""
        g, e = self.geo, self.emit
        f = self.fid[info.name]
        V, R = g.V, g.R
        if info.consumed and info.bc is not None:
            e("{")
            e("const bool pin = (unsigned)({}) < {}u;".format(plane_expr, self.NS), 3)
            e("if (!(pin && interior)) {", 3)
            e("const u32 m = pin ? cmask : 0u;", 4)
            e("#pragma unroll", 4)
            e("for (int r = 0; r < {}; ++r)".format(R), 4)
            e("#pragma unroll", 5)
            e("for (int v = 0; v < {V}; ++v) if (!((m >> (r * {V} + v)) & 1u)) nv[r][v] = {bc};".format(
                V=V, bc=self.lit(info.bc)), 5)
            e("}", 3)
            e("}")
        if not info.consumed:
            return
        for wdx in range(info.window - 1, 0, -1):
            e("#pragma unroll", 2)
            e("for (int r = 0; r < {}; ++r)".format(R), 2)
            e("#pragma unroll", 3)
            e("for (int v = 0; v < {V}; ++v) w_{f}[{a}][r][v] = w_{f}[{b}][r][v];".format(
                V=V, f=f, a=wdx, b=wdx - 1), 3)
        e("#pragma unroll", 2)
        e("for (int r = 0; r < {}; ++r)".format(R), 2)
        e("#pragma unroll", 3)
        e("for (int v = 0; v < {V}; ++v) w_{f}[0][r][v] = nv[r][v];".format(V=V, f=f), 3)
        if info.row_ring:
            n = info.row_reach
            # layout [ring][WR][2][n][TC]
            base = "xrow_{f} + ((xr_{f} * {WR} + wr) * 2) * {sz}".format(f=f, WR=g.WR, sz=n * g.TC)
            for q in range(n):
                e("sf_stv<{T}, {V}>({base} + {o} + c0, nv[{r}]);".format(
                    T=self.T, V=V, base=base, o=q * g.TC, r=q), 2)
                e("sf_stv<{T}, {V}>({base} + {o} + c0, nv[{r}]);".format(
                    T=self.T, V=V, base=base, o=(n + q) * g.TC, r=R - n + q), 2)
        if info.col_ring:
            # layout [ring][WR][WC][2][R][V]
            base = "xcol_{f} + (((xc_{f} * {WR} + wr) * {WC} + wc) * 2) * {sz}".format(
                f=f, WR=g.WR, WC=g.WC, sz=R * V)
            e("if (lane == 0) {", 2)
            for r in range(R):
                e("sf_stv<{T}, {V}>({base} + {o}, nv[{r}]);".format(T=self.T, V=V, base=base, o=r * V, r=r), 3)
            e("}", 2)
            e("if (lane == 31) {", 2)
            for r in range(R):
                e("sf_stv<{T}, {V}>({base} + {o}, nv[{r}]);".format(T=self.T, V=V, base=base, o=(R + r) * V, r=r), 3)
            e("}", 2)

    def _produce_ext(self, info: _FieldInfo):
        g, e = self.geo, self.emit
        f = self.fid[info.name]
        e("{  // input field " + f, 2)
        e("{T} nv[{R}][{V}];".format(T=self.T, R=g.R, V=g.V), 3)
        for r in range(g.R):
            row = "(r0 + {})".format(r) if self.ana.ndim == 3 else "0"
            e("sf_ldv<{T}, {V}>(nv[{r}], tile_{f} + slot * {sz} + {row} * {TC} + c0);".format(
                T=self.T, V=g.V, r=r, f=f, sz=g.TR * g.TC, row=row, TC=g.TC), 3)
        self._finish_field(info, "t - ({})".format(info.lag))
        e("}", 2)

    def _produce_op(self, op: StencilOp):
        g, a, e = self.geo, self.ana, self.emit
        info = a.fields[op.name]
        V, R, T = g.V, g.R, self.T
        taps = a.taps[op.name]
        plane = "t - ({})".format(info.lag)
        e("{  // operator producing " + self.fid[op.name], 2)
        # row vectors needed: (field, age, row) ; shifts needed per row vector
        rows = collections.OrderedDict()
        for (field, d, dj, dk) in taps:
            src = a.fields[field]
            age = info.lag - d - src.lag
            for r in range(R):
                key = (field, age, r + dj)
                ent = rows.setdefault(key, [0, 0])
                if dk < 0:
                    ent[0] = max(ent[0], -dk)
                if dk > 0:
                    ent[1] = max(ent[1], dk)
        names = {}
        for (field, age, rr), (nl, nr) in rows.items():
            src = a.fields[field]
            f = self.fid[field]
            tag = "{}_{}_{}".format(f, age, _fmt_off(rr))
            if 0 <= rr < R:
                vec = "w_{}[{}][{}]".format(f, age, rr)
                names[(field, age, rr)] = (vec, tag, "own")
                if nl or nr:
                    self._emit_shifts(src, vec, tag, rr, age, nl, nr)
            else:
                # a row owned by the neighbouring warp: read it (and its shifted columns) from the ring
                n = src.row_reach
                slot = self._slot_expr("xr_" + f, src.row_ring, age)
                if rr < 0:
                    nbr = "max(wr - 1, 0)"
                    side_row = n + (n + rr)          # bottom rows of the warp above
                else:
                    nbr = "min(wr + 1, {})".format(g.WR - 1)
                    side_row = rr - R                # top rows of the warp below
                base = "xrow_{f} + (({slot} * {WR} + {nbr}) * 2) * {sz} + {o}".format(
                    f=f, slot=slot, WR=g.WR, nbr=nbr, sz=n * g.TC, o=side_row * g.TC)
                e("const {T}* const p_{tag} = {base};".format(T=T, tag=tag, base=base), 3)
                e("{T} x_{tag}[{V}];".format(T=T, tag=tag, V=V), 3)
                e("sf_ldv<{T}, {V}>(x_{tag}, p_{tag} + c0);".format(T=T, V=V, tag=tag), 3)
                names[(field, age, rr)] = ("x_" + tag, tag, "ring")
                if nl:
                    e("{T} l_{tag}[{n}];".format(T=T, tag=tag, n=nl), 3)
                    for q in range(nl):
                        e("l_{tag}[{q}] = p_{tag}[max(c0 - {nl} + {q}, 0)];".format(tag=tag, q=q, nl=nl), 3)
                if nr:
                    e("{T} g_{tag}[{n}];".format(T=T, tag=tag, n=nr), 3)
                    for q in range(nr):
                        e("g_{tag}[{q}] = p_{tag}[min(c0 + {V} + {q}, {last})];".format(
                            tag=tag, q=q, V=V, last=g.TC - 1), 3)

        def tap_c(t: ex.Tap, r: int, v: int) -> str:
            if a.ndim == 3:
                d, dj, dk = t.offset
            else:
                d, dj, dk = t.offset[1], 0, t.offset[2]
            src = a.fields[t.field]
            age = info.lag - d - src.lag
            vec, tag, kind = names[(t.field, age, r + dj)]
            c = v + dk
            nl, nr = rows[(t.field, age, r + dj)]
            if 0 <= c < V:
                return "{}[{}]".format(vec, c)
            if c < 0:
                return "l_{}[{}]".format(tag, nl + c)
            return "g_{}[{}]".format(tag, c - V)

        math_fn = _MATH_F32 if self.ct == dtypes.float32 else _MATH_F64
        e("{T} nv[{R}][{V}];".format(T=T, R=R, V=V), 3)
        local_ids = {}
        for r in range(R):
            for v in range(V):
                local = {}
                for s in op.statements:
                    rhs = ex.emit_c(
                        s.value,
                        tap=lambda t, r=r, v=v: tap_c(t, r, v),
                        var=lambda n, local=local: self._var(n, local, op),
                        literal=self.lit,
                        call=lambda fn, args: "{}({})".format(math_fn[fn], ", ".join(args)))
                    if s.target not in local_ids:
                        local_ids[s.target] = len(local_ids)
                    lname = "q{}_{}_{}".format(local_ids[s.target], r, v)
                    ty = "const bool" if isinstance(s.value, (ex.Cmp, ex.Logic)) else "const " + T
                    if s.target in local:
                        lname += "b"
                    e("{} {} = {};".format(ty, lname, rhs), 3)
                    local[s.target] = lname
                target = op.name if op.name in local else op.statements[-1].target
                e("nv[{}][{}] = {};".format(r, v, local[target]), 3)
        if info.stored:
            idx = [i.name for i in a.fields.values() if i.stored].index(op.name)
            e("if (({p}) >= c_begin && ({p}) < c_end) {{".format(p=plane), 3)
            if a.ndim == 3:
                e("{T}* const op = o_{n} + ((i64)(({p}) - s_base) * {NJ} + gj0) * {NK} + gk;".format(
                    T=T, n=idx, p=plane, NJ=self.NJ, NK=self.NK), 4)
            else:
                e("{T}* const op = o_{n} + (i64)(({p}) - s_base) * {NK} + gk;".format(
                    T=T, n=idx, p=plane, NK=self.NK), 4)
            for r in range(R):
                e("if (smask & {m}u) sf_stv<{T}, {V}>(op + {o}, nv[{r}]);".format(
                    m=1 << r, T=T, V=V, o=r * self.NK, r=r), 4)
            e("}", 3)
        self._finish_field(info, plane)
        e("}", 2)

    def _emit_shifts(self, src, vec, tag, rr, age, nl, nr):
        """Left/right neighbour cells of an owned row: warp shuffles, plus the column ring at warp
        edges when several warps share a row."""
        g, e, T, V = self.geo, self.emit, self.T, self.geo.V
        f = self.fid[src.name]
        if nl:
            e("{T} l_{tag}[{n}];".format(T=T, tag=tag, n=nl), 3)
            for q in range(nl):
                e("l_{tag}[{q}] = __shfl_up_sync(0xffffffffu, {vec}[{c}], 1);".format(
                    tag=tag, q=q, vec=vec, c=V - nl + q), 3)
        if nr:
            e("{T} g_{tag}[{n}];".format(T=T, tag=tag, n=nr), 3)
            for q in range(nr):
                e("g_{tag}[{q}] = __shfl_down_sync(0xffffffffu, {vec}[{c}], 1);".format(
                    tag=tag, q=q, vec=vec, c=q), 3)
        if src.col_ring and (nl or nr):
            slot = self._slot_expr("xc_" + f, src.col_ring, age)
            sz = g.R * V
            if nl:
                e("if (lane == 0 && wc > 0) {", 3)
                base = "xcol_{f} + ((({slot} * {WR} + wr) * {WC} + wc - 1) * 2 + 1) * {sz} + {o}".format(
                    f=f, slot=slot, WR=g.WR, WC=g.WC, sz=sz, o=rr * V)
                for q in range(nl):
                    e("l_{tag}[{q}] = ({base})[{c}];".format(tag=tag, q=q, base=base, c=V - nl + q), 4)
                e("}", 3)
            if nr:
                e("if (lane == 31 && wc < {}) {{".format(g.WC - 1), 3)
                base = "xcol_{f} + ((({slot} * {WR} + wr) * {WC} + wc + 1) * 2) * {sz} + {o}".format(
                    f=f, slot=slot, WR=g.WR, WC=g.WC, sz=sz, o=rr * V)
                for q in range(nr):
                    e("g_{tag}[{q}] = ({base})[{c}];".format(tag=tag, q=q, base=base, c=q), 4)
                e("}", 3)

    def _var(self, name, local, op):
        if name in local:
            return local[name]
        if name in self.specialize:
            return self.lit(self.specialize[name])
        if name in self.program.constants:
            return self.lit(self.program.constants[name]["value"])
        if name in self.sc_name:
            return "({}){}".format(self.T, self.sc_name[name])
        raise NameError("Unknown name {} in operator {}".format(name, op.name))


# ---------------------------------------------------------------------------------- planning


def _vector_width(program, dtype):
    v = 16 // dtype.bytes
    return v if program.shape[-1] % v == 0 else None


def choose_geometry(program, ops, options) -> Optional[Tuple[GroupAnalysis, Geometry]]:
    """Pick (V, R, warps, prefetch) for a candidate group or return None when it cannot stream."""
    dtype = ops[0].data_type
    ndim = len(program.shape)
    try:
        V = _vector_width(program, dtype)
        if V is None:
            raise NotStreamable("innermost extent not a multiple of the vector width")
        prefetch = options.prefetch or 2
        candidates = []
        nk = program.shape[-1]
        if ndim == 3:
            warps = options.warps or 16
            rows = [options.rows_per_thread] if options.rows_per_thread else [4, 3, 2, 1]
            for R in rows:
                candidates.append((R, warps, 1, 32))
            if nk < 32 * V and nk // V <= 64:
                # narrow innermost dimension: one row of nk/V threads, several row groups per warp
                ks = nk // V
                for R in rows:
                    for groups in (options.warps * 32 // ks if options.warps else 0, 32, 24, 16):
                        if groups and (groups * ks) % 32 == 0 and groups * ks <= 768:
                            candidates.append((R, groups, 1, ks))
        else:
            warps = options.warps or 8
            candidates.append((1, 1, warps, 32))
        best = None
        for (R, WR, WC, KS) in candidates:
            try:
                ana = GroupAnalysis(program, ops, exchange_cols=(WC > 1))
                geo = Geometry(ana, V, R, WR, WC, prefetch, KS)
            except NotStreamable:
                continue
            if geo.smem > SMEM_LIMIT:
                continue
            if ana.window_registers(R, V) > REG_BUDGET:
                continue
            eff = (geo.BJ * geo.BK) / float(geo.TR * geo.TC) if ndim == 3 else geo.BK / float(geo.TC)
            eff *= min(1.0, nk / float(geo.BK))          # lanes beyond a narrow domain are idle
            if best is None or eff > best[0] + 1e-9:
                best = (eff, ana, geo)
        if best is None:
            return None
        return best[1], best[2]
    except NotStreamable:
        return None


# Cost model of the planner (seconds).  Calibrated on B200 (profiles/sweep_depth_rows_r01.txt):
# a streamed pass costs max(HBM time of its algorithmic bytes, issue time of the cells it computes
# including the redundant halo cells).
HBM_BYTES_PER_S = 6.4e12
UPDATES_PER_S = {4: 1.9e12, 8: 0.9e12}      # computed cell updates per second, by element size
GENERAL_EFFICIENCY = 0.84                   # fraction of HBM bandwidth the one-operator kernel reaches


def _tile_efficiency(ana, geo):
    if ana.ndim == 3:
        return (geo.BJ * geo.BK) / float(geo.TR * geo.TC)
    return geo.BK / float(geo.TC)


def group_cost(program, ops, options):
    """Estimated time of one streamed pass over ``ops`` (None if the group cannot stream)."""
    chosen = choose_geometry(program, ops, options)
    if chosen is None:
        return None
    ana, geo = chosen
    fields = program.fields
    nbytes = sum(fields[i.name].nbytes for i in ana.ext_fields)
    nbytes += sum(fields[i.name].nbytes for i in ana.fields.values() if i.stored)
    eff = _tile_efficiency(ana, geo)
    used = min(1.0, program.shape[-1] / float(geo.BK))
    t_mem = nbytes / HBM_BYTES_PER_S
    t_cmp = program.cells * len(ops) / (eff * used) / UPDATES_PER_S[ana.dtype.bytes]
    return max(t_mem, t_cmp) + 5e-6


def general_cost(program, op):
    fields = program.fields
    nbytes = fields[op.name].nbytes + sum(fields[f].nbytes for f in op.accesses)
    return nbytes / (HBM_BYTES_PER_S * GENERAL_EFFICIENCY) + 5e-6


def partition(program: StencilProgram, options):
    """Cut the topologically ordered operators into passes minimising the modelled run time
    (dynamic programme over contiguous groups of at most ``max_depth`` operators)."""
    ops = list(program.ops)
    n = len(ops)
    max_depth = options.max_depth or 8
    best = [0.0] + [None] * n          # best[k]: cost of the first k operators
    choice = [None] * (n + 1)
    cache = {}
    for end in range(1, n + 1):
        for start in range(max(0, end - max_depth), end):
            group = ops[start:end]
            # structurally identical chain segments (Jacobi chains) share one estimate
            chain = _chain_like(group)
            key = tuple((repr([op.offsets3(f) for f in op.accesses]),
                         repr(sorted(op.boundary_conditions.items(), key=repr)[0][1:] if op.boundary_conditions else ""),
                         op.data_type.name) for op in group) if chain else None
            if chain and key in cache:
                cost = cache[key]
            else:
                cost = group_cost(program, group, options)
                if chain:
                    cache[key] = cost
            family = "streamed"
            if cost is None:
                if len(group) > 1:
                    continue
                cost, family = general_cost(program, group[0]), "general"
            if best[start] is None:
                continue
            total = best[start] + cost
            if best[end] is None or total < best[end] - 1e-12:
                best[end] = total
                choice[end] = (start, family)
    groups = []
    end = n
    while end > 0:
        start, family = choice[end]
        groups.append((family, ops[start:end]))
        end = start
    groups.reverse()
    return groups


def _chain_like(group):
    """True when every operator reads exactly one array field: the previous operator's result
    (or the group's single input) -- the case where cost estimates can be shared."""
    prev = None
    for op in group:
        if len(op.accesses) != 1:
            return False
        (field,) = op.accesses.keys()
        if prev is not None and field != prev:
            return False
        prev = op.name
    return True


def choose_chunk(n_stream, tiles, overhead, sms=148):
    """Planes per CTA along the streamed dimension: balance wave quantisation against the
    redundant warm-up planes every chunk recomputes."""
    best = None
    for chunks in range(1, 65):
        ci = -(-n_stream // chunks)
        blocks = tiles * (-(-n_stream // ci))
        waves = -(-blocks // sms)
        cost = waves * (ci + overhead)
        if best is None or cost < best[0]:
            best = (cost, ci)
    return best[1]


def lower_group(lowered: LoweredProgram, ops: List[StencilOp], options, specialize=None):
    program = lowered.program
    chosen = choose_geometry(program, ops, options)
    if chosen is None:
        raise NotStreamable("group cannot stream")
    ana, geo = chosen
    gen = StreamKernelGen(program, ops, ana, geo, specialize)
    name, src, args = gen.generate()
    if name not in lowered.kernels:
        lowered.kernels[name] = KernelSpec(name, src, (geo.NT, 1, 1), "streamed")
    NI, NJ, NK = program.shape3
    n_stream = program.shape[0]
    gx = -(-NK // geo.BK)
    gy = -(-NJ // geo.BJ) if ana.ndim == 3 else 1
    overhead = ana.t_end_offset() - ana.t_begin_offset()

    def chunk_for(b, e_):
        if options.chunk:
            return options.chunk
        return choose_chunk(max(1, e_ - b), gx * gy, overhead)

    def grid(b, e_):
        ci = chunk_for(b, e_)
        return (gx, gy, max(1, -(-(e_ - b) // ci)))

    stored = [i.name for i in ana.fields.values() if i.stored]
    reads = [i.name for i in ana.ext_fields]
    launch = LaunchSpec(kernel=name, grid_fn=grid, block=(geo.NT, 1, 1), smem=geo.smem, args=args,
                        ops=[op.name for op in ops], reads=reads, writes=stored,
                        cells_per_unit=program.cells * len(ops), family="streamed",
                        info={"V": geo.V, "R": geo.R, "warps": [geo.WR, geo.WC], "threads_per_row": geo.KS,
                              "tile": [geo.TR, geo.TC],
                              "block_out": [geo.BJ, geo.BK], "halo": [geo.HJ0, geo.HJ1, geo.HK0, geo.HK1],
                              "prefetch": geo.P, "lags": {n: i.lag for n, i in ana.fields.items()},
                              "windows": {n: i.window for n, i in ana.fields.items() if i.consumed},
                              "window_registers": ana.window_registers(geo.R, geo.V),
                              "stream_overhead_planes": overhead,
                              "back": max(i.back for i in ana.fields.values()),
                              "reach": {i.name: (i.back, i.fwd) for i in ana.ext_fields},
                              "tile_efficiency": _tile_efficiency(ana, geo),
                              "fwd": ana.max_lag,
                              "chunk_fn": chunk_for})
    lowered.launches.append(launch)
    lowered.uses_stream = True
    return launch

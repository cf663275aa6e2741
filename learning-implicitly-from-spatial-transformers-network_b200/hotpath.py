"""Host side of the hot path: thin, torch-typed wrappers over the C ABI (include/list_b200.h).

torch provides device memory (caching allocator), streams and autograd plumbing only; every
arithmetic step of rows a-1..a-9 runs in liblist_b200.so.  Nothing here falls back
to PyTorch ops: inputs that are not CUDA tensors raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _C

MAP_SIZE = 137           # reference network/modules.py:16
DEFAULT_CHUNK = 65536    # reference arguments.py:18 (test_pointnum)

_TORCH_DTYPE = {_C.F32: torch.float32, _C.BF16: torch.bfloat16}
_DTYPE_CODE = {torch.float32: _C.F32, torch.bfloat16: _C.BF16, "fp32": _C.F32, "bf16": _C.BF16,
               "float32": _C.F32, "bfloat16": _C.BF16}
_device_checked = False


def dtype_code(dtype) -> int:
    try:
        return _DTYPE_CODE[dtype]
    except KeyError:
        raise ValueError(f"unsupported dtype {dtype!r}: the hot path computes in fp32 or bf16") from None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*tensors) -> torch.device:
    global _device_checked
    dev = None
    for t in tensors:
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError("list_b200: tensors must live on a CUDA device (there is no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"list_b200: tensors on different devices ({dev} vs {t.device})")
    if not _device_checked:
        with torch.cuda.device(dev):
            _C.check(_C.lib().list_b200_device_ok(), "list_b200_device_ok")
        _device_checked = True
    return dev


# ------------------------------------------------------------------------------ layout
@dataclass(frozen=True)
class FeatureLayout:
    k_out: int
    k_pad: int
    map_off: int
    xyz_off: int
    vol_off: tuple
    perm: np.ndarray          # perm[new_col] = reference column (modules.py:270-275 order)


def feature_layout(map_channels: int, vol_ch: Sequence[int]) -> FeatureLayout:
    lay = _C.ListLayout()
    k_out = map_channels + _C.NUM_DISP * sum(vol_ch) + 3
    perm = (C.c_int32 * k_out)()
    _C.check(_C.lib().list_feature_layout(map_channels, len(vol_ch), _C.i32_array(vol_ch), C.byref(lay), perm),
             "list_feature_layout")
    assert lay.k_out == k_out
    return FeatureLayout(lay.k_out, lay.k_pad, lay.map_off, lay.xyz_off, tuple(lay.vol_off[:len(vol_ch)]),
                         np.ctypeslib.as_array(perm).copy())


# ------------------------------------------------------------------------------ per-image context
@dataclass
class HotPathContext:
    """Per-image tensors in the kernels' channels-last layout (C struct ListCtx)."""
    maps_cl: torch.Tensor               # (B, S, S, Cm)
    vols_cl: List[torch.Tensor]         # (B, R, R, R, C) each
    trans_mat: torch.Tensor             # (B, 4, 3) fp32
    dtype: int

    @property
    def B(self) -> int:
        return self.maps_cl.shape[0]

    @property
    def vol_ch(self) -> List[int]:
        return [v.shape[-1] for v in self.vols_cl]

    @property
    def layout(self) -> FeatureLayout:
        return feature_layout(self.maps_cl.shape[-1], self.vol_ch)

    def struct(self) -> _C.ListCtx:
        s = _C.ListCtx()
        s.B = self.B
        s.dtype = self.dtype
        s.map_size = self.maps_cl.shape[1]
        s.map_channels = self.maps_cl.shape[3]
        s.maps = self.maps_cl.data_ptr()
        s.n_levels = len(self.vols_cl)
        for l, v in enumerate(self.vols_cl):
            s.vol_res[l] = v.shape[1]
            s.vol_ch[l] = v.shape[4]
            s.vols[l] = v.data_ptr()
        s.trans_mat = self.trans_mat.data_ptr()
        return s


def prepare_context(maps: Sequence[torch.Tensor], vols: Sequence[torch.Tensor], trans_mat: torch.Tensor,
                    dtype="fp32", map_size: int = MAP_SIZE, skip_levels: Sequence[int] = ()) -> HotPathContext:
    """Once per image (hoists reference modules.py:25-35 out of the chunk loop): NCHW maps ->
    upsampled channels-last, NCDHW volumes -> channels-last, optional bf16.  Levels in `skip_levels` only get their
    buffer; grid_sdf_late fills them (staged evaluation)."""
    code = dtype_code(dtype)
    dev = _require_cuda(*maps, *vols, trans_mat)
    lib = _C.lib()
    maps = [m.detach().to(torch.float32).contiguous() for m in maps]
    vols = [v.detach().to(torch.float32).contiguous() for v in vols]
    B = maps[0].shape[0]
    for m in maps:
        if m.dim() != 4 or m.shape[0] != B or m.shape[2] != m.shape[3]:
            raise ValueError(f"feature map of shape {tuple(m.shape)}: expected (B, C, H, H)")
    for v in vols:
        if v.dim() != 5 or v.shape[0] != B or not (v.shape[2] == v.shape[3] == v.shape[4]):
            raise ValueError(f"volume of shape {tuple(v.shape)}: expected (B, C, R, R, R)")
    if tuple(trans_mat.shape) != (B, 4, 3):
        raise ValueError(f"trans_mat of shape {tuple(trans_mat.shape)}: expected ({B}, 4, 3)")
    cm = sum(m.shape[1] for m in maps)
    tdt = _TORCH_DTYPE[code]
    with torch.cuda.device(dev):
        maps_cl = torch.empty(B, map_size, map_size, cm, device=dev, dtype=tdt)
        _C.check(lib.list_prep_maps(_C.ptr_array([m.data_ptr() for m in maps]),
                                    _C.i32_array([m.shape[1] for m in maps]),
                                    _C.i32_array([m.shape[2] for m in maps]),
                                    len(maps), B, map_size, maps_cl.data_ptr(), code, _stream()), "list_prep_maps")
        vols_cl = []
        for l, v in enumerate(vols):
            out = torch.empty(B, v.shape[2], v.shape[3], v.shape[4], v.shape[1], device=dev, dtype=tdt)
            if l not in skip_levels:
                _C.check(lib.list_prep_volume(v.data_ptr(), B, v.shape[1], v.shape[2], out.data_ptr(), code, _stream()),
                         "list_prep_volume")
            vols_cl.append(out)
    return HotPathContext(maps_cl, vols_cl, trans_mat.detach().to(torch.float32).contiguous(), code)


# ------------------------------------------------------------------------------ weights
@dataclass
class KernelWeights:
    """Kernel-format copies of sdf_decoder.fc.* (C struct ListWeights).  Derived buffers: the
    nn.Parameters under the reference's state_dict keys stay the masters."""
    w0: torch.Tensor      # (n0, k_pad) columns permuted to the gather's order, zero padded
    w1: torch.Tensor
    w2: torch.Tensor
    w3: torch.Tensor      # (n2,) fp32
    b0: torch.Tensor
    b1: torch.Tensor
    b2: torch.Tensor
    b3: torch.Tensor
    dtype: int
    k_pad: int

    def struct(self) -> _C.ListWeights:
        s = _C.ListWeights()
        s.dtype = self.dtype
        s.k_pad = self.k_pad
        s.n0, s.n1, s.n2 = self.w0.shape[0], self.w1.shape[0], self.w2.shape[0]
        for name in ("w0", "w1", "w2", "w3", "b0", "b1", "b2", "b3"):
            setattr(s, name, getattr(self, name).data_ptr())
        return s


def prepare_weights(state: dict, layout: FeatureLayout, dtype="fp32", prefix: str = "fc.") -> KernelWeights:
    """`state` holds the reference's Conv1d tensors ('fc.fc_0.weight' [512,3610,1], ... modules.py:196-200)."""
    code = dtype_code(dtype)
    w0 = state[f"{prefix}fc_0.weight"]
    dev = _require_cuda(w0)
    tdt = _TORCH_DTYPE[code]
    w0 = w0.detach().reshape(w0.shape[0], -1).to(torch.float32)
    if w0.shape[1] != layout.k_out:
        raise ValueError(f"fc_0 has {w0.shape[1]} input features, the layout expects {layout.k_out}")
    perm = torch.from_numpy(layout.perm.astype(np.int64)).to(dev)
    w0p = torch.zeros(w0.shape[0], layout.k_pad, device=dev, dtype=torch.float32)
    w0p[:, :layout.k_out] = w0[:, perm]

    def mat(name):
        w = state[f"{prefix}{name}.weight"].detach()
        return w.reshape(w.shape[0], -1).to(torch.float32)

    def vec(name):
        return state[f"{prefix}{name}.bias"].detach().to(torch.float32).contiguous()

    return KernelWeights(
        w0=w0p.to(tdt).contiguous(), w1=mat("fc_1").to(tdt).contiguous(), w2=mat("fc_2").to(tdt).contiguous(),
        w3=mat("fc_out").reshape(-1).contiguous(), b0=vec("fc_0"), b1=vec("fc_1"), b2=vec("fc_2"), b3=vec("fc_out"),
        dtype=code, k_pad=layout.k_pad)


# ------------------------------------------------------------------------------ forward pieces
def grid_points(res: int, begin: int = 0, count: Optional[int] = None, bb_min: float = -0.5, bb_max: float = 0.5,
                device="cuda") -> torch.Tensor:
    """utils.create_grid_points_from_bounds rows [begin, begin+count) as fp32 (count, 3)."""
    count = res ** 3 - begin if count is None else count
    q = torch.empty(count, 3, device=device, dtype=torch.float32)
    _require_cuda(q)
    with torch.cuda.device(q.device):
        _C.check(_C.lib().list_grid_points(q.data_ptr(), res, bb_min, bb_max, begin, count, _stream()), "list_grid_points")
    return q


def gather_features(ctx: HotPathContext, points: torch.Tensor, raw: bool = True) -> torch.Tensor:
    """Rows a-2..a-5: (B, N, 3) query points -> feature rows X (B*N, k_pad) of ctx.dtype."""
    dev = _require_cuda(ctx.maps_cl, points)
    B, N, _ = points.shape
    pts = points.detach().to(torch.float32).contiguous()
    lay = ctx.layout
    X = torch.empty(B * N, lay.k_pad, device=dev, dtype=_TORCH_DTYPE[ctx.dtype])
    cs = ctx.struct()
    with torch.cuda.device(dev):
        _C.check(_C.lib().list_gather_fwd(C.byref(cs), pts.data_ptr(), int(raw), X.data_ptr(), lay.k_pad, B, N, _stream()),
                 "list_gather_fwd")
    return X


def gather_grid_features(ctx: HotPathContext, image: int, res: int, begin: int, count: int,
                         bb_min: float = -0.5, bb_max: float = 0.5) -> torch.Tensor:
    dev = _require_cuda(ctx.maps_cl)
    lay = ctx.layout
    X = torch.empty(count, lay.k_pad, device=dev, dtype=_TORCH_DTYPE[ctx.dtype])
    cs = ctx.struct()
    with torch.cuda.device(dev):
        _C.check(_C.lib().list_gather_grid_fwd(C.byref(cs), image, res, bb_min, bb_max, begin, count, X.data_ptr(),
                                               lay.k_pad, _stream()), "list_gather_grid_fwd")
    return X


# ------------------------------------------------------------------------------ hoisted fc_0 (dense grids, bf16)
class HoistedState:
    """Output of list_hoist_prepare: the caller-owned buffer with the projected maps / coarse levels
    (csrc/hoist.cu) of every image of `ctx`, for `weights`."""

    def __init__(self, ctx: HotPathContext, weights: KernelWeights):
        dev = _require_cuda(ctx.maps_cl, weights.w0)
        lib = _C.lib()
        self.ctx, self.base = ctx, weights
        cs, ws = ctx.struct(), weights.struct()
        need = lib.list_hoist_bytes(C.byref(cs), C.byref(ws))
        if need == 0:
            raise RuntimeError("list_hoist_bytes returned 0: this configuration has no hoisted path")
        hc, kh = C.c_int32(0), C.c_int32(0)
        _C.check(lib.list_hoist_layout(C.byref(cs), C.byref(ws), C.byref(hc), C.byref(kh)), "list_hoist_layout")
        self.hoist_cols, self.k_h = hc.value, kh.value
        self.buf = torch.empty(need, device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            _C.check(lib.list_hoist_prepare(C.byref(cs), C.byref(ws), self.buf.data_ptr(), need, _stream()),
                     "list_hoist_prepare")

    def mlp(self, X: torch.Tensor, out_div: float = 1.0) -> torch.Tensor:
        """Row a-6 on hoisted rows X (rows, k_h) = [addend 512 | remaining columns]."""
        dev = _require_cuda(X, self.base.w0)
        if X.dtype != torch.bfloat16 or X.shape[1] < self.k_h or X.stride(1) != 1:
            raise ValueError("hoisted rows must be bf16 with at least k_h contiguous columns")
        rows = X.shape[0]
        sdf = torch.empty(rows, device=dev, dtype=torch.float32)
        ws = self.base.struct()
        with torch.cuda.device(dev):
            _C.check(_C.lib().list_mlp_hoisted_fwd(C.byref(ws), self.hoist_cols, X.data_ptr(), X.stride(0), rows,
                                                   sdf.data_ptr(), float(out_div), _stream()), "list_mlp_hoisted_fwd")
        return sdf

    def gather_grid(self, image: int, res: int, begin: int, count: int, bb_min: float = -0.5, bb_max: float = 0.5,
                    parts: int = 3, out: Optional[torch.Tensor] = None):
        """Hoisted feature rows; parts: 1 = addend columns only, 2 = remaining columns only, 3 = whole row."""
        dev = self.buf.device
        X = out if out is not None else torch.empty(count, self.k_h, device=dev, dtype=torch.bfloat16)
        cs, ws = self.ctx.struct(), self.base.struct()
        with torch.cuda.device(dev):
            _C.check(_C.lib().list_hoist_gather_grid_fwd(C.byref(cs), C.byref(ws), self.buf.data_ptr(), image, res, bb_min,
                                                         bb_max, begin, count, X.data_ptr(), self.k_h, parts, _stream()),
                     "list_hoist_gather_grid_fwd")
        return X


class LineTableState:
    """The stages of the line-table dense-grid path (csrc/lines.cu + csrc/grid_tc.cu; what list_sdf_grid runs for bf16)
    exposed one by one for tests and per-kernel timing: projection once per image, then per range of grid points the
    per-line column tables G, the non-hoisted feature columns Xr and the fused interpolation + MLP kernel."""

    def __init__(self, ctx: HotPathContext, weights: KernelWeights):
        dev = _require_cuda(ctx.maps_cl, weights.w0)
        lib = _C.lib()
        self.ctx, self.base, self.dev = ctx, weights, dev
        cs, ws = ctx.struct(), weights.struct()
        need = lib.list_lines_hoist_bytes(C.byref(cs), C.byref(ws))
        if need == 0:
            raise RuntimeError("list_lines_hoist_bytes returned 0: this configuration has no line-table path")
        hc, kf, rpl = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        _C.check(lib.list_lines_layout(C.byref(cs), C.byref(ws), C.byref(hc), C.byref(kf), C.byref(rpl)), "list_lines_layout")
        self.hoist_cols, self.k_f, self.rows_per_line = hc.value, kf.value, rpl.value
        self.buf = torch.empty(need, device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            _C.check(lib.list_lines_prepare(C.byref(cs), C.byref(ws), self.buf.data_ptr(), need, _stream()), "list_lines_prepare")

    def table(self, image: int, res: int, begin: int, count: int, bb_min: float = -0.5, bb_max: float = 0.5,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """G (lines, rows_per_line, 512) bf16 for the z-lines touched by grid points [begin, begin+count)."""
        cs, ws = self.ctx.struct(), self.base.struct()
        lib = _C.lib()
        need = lib.list_lines_table_bytes(C.byref(cs), C.byref(ws), res, begin, count)
        lines = need // (self.rows_per_line * 1024)
        G = out if out is not None else torch.empty(lines, self.rows_per_line, 512, device=self.dev, dtype=torch.bfloat16)
        with torch.cuda.device(self.dev):
            _C.check(lib.list_lines_table(C.byref(cs), C.byref(ws), self.buf.data_ptr(), image, res, bb_min, bb_max, begin, count,
                                          G.data_ptr(), G.numel() * 2, _stream()), "list_lines_table")
        return G

    def rest(self, image: int, res: int, begin: int, count: int, bb_min: float = -0.5, bb_max: float = 0.5,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Xr (count, k_f) bf16: the feature columns that are not hoisted (plus 1.0 in the three pad columns behind q,
        through which fc_0's bias enters the MMA)."""
        X = out if out is not None else torch.empty(count, self.k_f, device=self.dev, dtype=torch.bfloat16)
        cs, ws = self.ctx.struct(), self.base.struct()
        with torch.cuda.device(self.dev):
            _C.check(_C.lib().list_lines_rest(C.byref(cs), C.byref(ws), image, res, bb_min, bb_max, begin, count, X.data_ptr(),
                                              X.stride(0), _stream()), "list_lines_rest")
        return X

    def plan(self, image: int, res: int, begin: int, count: int, G: torch.Tensor, bb_min: float = -0.5, bb_max: float = 0.5,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Tile plans (row lists pointing into self.buf and G, per-step interpolation weights) as an opaque uint8 buffer."""
        lib = _C.lib()
        need = lib.list_grid_plan_bytes(res, begin, count)
        buf = out if out is not None else torch.empty(need, device=self.dev, dtype=torch.uint8)
        cs, ws = self.ctx.struct(), self.base.struct()
        with torch.cuda.device(self.dev):
            _C.check(lib.list_grid_plan(C.byref(cs), C.byref(ws), self.buf.data_ptr(), image, res, bb_min, bb_max, begin, count,
                                        G.data_ptr(), buf.data_ptr(), buf.numel(), _stream()), "list_grid_plan")
        return buf

    def evaluate(self, res: int, begin: int, count: int, Xr: torch.Tensor, plan: torch.Tensor, out_div: float = 1.0,
                 bb_min: float = -0.5, bb_max: float = 0.5, debug: bool = False, trace: bool = False,
                 stats: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
        """Fused interpolation + MLP.  Returns sdf (count,) [, relu(fc_0) (count, 512) fp32] [, trace (16, 24) int64].
        stats: optional device int64[2] accumulating (tile pairs, executed 16-row k-steps of interpolation chunks).  The G tensor the plan was built
        for must still be alive."""
        sdf = out if out is not None else torch.empty(count, device=self.dev, dtype=torch.float32)
        h1 = torch.zeros(count, 512, device=self.dev, dtype=torch.float32) if debug else None
        tr = torch.zeros(16, 24, device=self.dev, dtype=torch.int64) if trace else None
        cs, ws = self.ctx.struct(), self.base.struct()
        with torch.cuda.device(self.dev):
            _C.check(_C.lib().list_grid_tc_fwd(C.byref(cs), C.byref(ws), self.buf.data_ptr(), res, bb_min, bb_max, begin, count,
                                               Xr.data_ptr(), Xr.stride(0), plan.data_ptr(), sdf.data_ptr(), float(out_div),
                                               None if h1 is None else h1.data_ptr(), None if tr is None else tr.data_ptr(),
                                               None if stats is None else stats.data_ptr(), _stream()), "list_grid_tc_fwd")
        out = [sdf]
        if debug:
            out.append(h1)
        if trace:
            out.append(tr)
        return out[0] if len(out) == 1 else tuple(out)


def mlp(weights: KernelWeights, X: torch.Tensor, out_div: float = 1.0, return_workspace: bool = False, train: bool = False):
    """Row a-6 on feature rows X (rows, ldx).  train=True (fp32 only): the forward of a training step (list_mlp_fwd_train),
    whose saved activations feed the backward."""
    dev = _require_cuda(X, weights.w0)
    if _DTYPE_CODE.get(X.dtype) != weights.dtype:
        raise ValueError(f"X dtype {X.dtype} does not match the weights' dtype code {weights.dtype}")
    rows = X.shape[0]
    sdf = torch.empty(rows, device=dev, dtype=torch.float32)
    ws_struct = weights.struct()
    lib = _C.lib()
    need = lib.list_mlp_workspace_bytes(C.byref(ws_struct), rows)
    ws = torch.empty(max(need, 16), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        fn = lib.list_mlp_fwd_train if train else lib.list_mlp_fwd
        _C.check(fn(C.byref(ws_struct), X.data_ptr(), X.stride(0), rows, sdf.data_ptr(), float(out_div),
                    ws.data_ptr(), ws.numel(), _stream()), "list_mlp_fwd")
    return (sdf, ws) if return_workspace else sdf


def _workspace(ctx_s, w_s, chunk_rows, dev, res: Optional[int] = None):
    """Workspace of list_sdf_fwd (res None: full feature rows) or of the dense-grid calls at resolution `res` (what the
    path that runs needs; much smaller for the bf16 line-table path)."""
    if res is None:
        need = _C.lib().list_sdf_workspace_bytes(C.byref(ctx_s), C.byref(w_s), chunk_rows)
    else:
        need = _C.lib().list_sdf_grid_workspace_bytes(C.byref(ctx_s), C.byref(w_s), res, chunk_rows)
    if need == 0:
        raise RuntimeError("workspace size query returned 0 (invalid ctx/weights)")
    return torch.empty(need, device=dev, dtype=torch.uint8)


def fit_grid_chunk(ctx: HotPathContext, weights: KernelWeights, res: int, chunk_rows: int, fraction: float = 0.6,
                   floor: int = 65536) -> int:
    """Largest chunk_rows <= the requested one whose dense-grid workspace fits into `fraction` of the device memory that is
    free right now (halving; never below `floor`)."""
    cs, wsn = ctx.struct(), weights.struct()
    free = torch.cuda.mem_get_info(ctx.maps_cl.device)[0]
    while chunk_rows > floor and _C.lib().list_sdf_grid_workspace_bytes(C.byref(cs), C.byref(wsn), res, chunk_rows) > fraction * free:
        chunk_rows //= 2
    return max(chunk_rows, 1)


def query_sdf(ctx: HotPathContext, weights: KernelWeights, points: torch.Tensor, raw: bool = True,
              out_div: float = 1.0, chunk_rows: int = DEFAULT_CHUNK) -> torch.Tensor:
    """Row a-7 (reference models.py:91-97): (B, N, 3) query points -> scaled SDF (B, N)."""
    dev = _require_cuda(ctx.maps_cl, weights.w0, points)
    B, N, _ = points.shape
    pts = points.detach().to(torch.float32).contiguous()
    sdf = torch.empty(B, N, device=dev, dtype=torch.float32)
    if N == 0:
        return sdf
    chunk_rows = max(1, min(chunk_rows, N))
    cs, wsn = ctx.struct(), weights.struct()
    ws = _workspace(cs, wsn, chunk_rows, dev)
    with torch.cuda.device(dev):
        _C.check(_C.lib().list_sdf_fwd(C.byref(cs), C.byref(wsn), pts.data_ptr(), int(raw), B, N, sdf.data_ptr(),
                                       float(out_div), chunk_rows, ws.data_ptr(), ws.numel(), _stream()), "list_sdf_fwd")
    return sdf


def grid_sdf(ctx: HotPathContext, weights: KernelWeights, res: int, begin: int = 0, count: Optional[int] = None,
             sdf_scale: float = 1.0, chunk_rows: int = DEFAULT_CHUNK, bb_min: float = -0.5, bb_max: float = 0.5,
             out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Row a-8 (reference executors.py:191-231): SDF/sdf_scale of grid points [begin, begin+count)
    for every image -> (B, count).  A rank's shard of the dense grid."""
    dev = _require_cuda(ctx.maps_cl, weights.w0)
    count = res ** 3 - begin if count is None else count
    if out is None:
        out = torch.empty(ctx.B, count, device=dev, dtype=torch.float32)
    elif out.numel() != ctx.B * count or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous fp32 tensor of B*count elements")
    if count == 0:
        return out
    chunk_rows = max(1, min(chunk_rows, count))
    cs, wsn = ctx.struct(), weights.struct()
    ws = workspace if workspace is not None else _workspace(cs, wsn, chunk_rows, dev, res)
    with torch.cuda.device(dev):
        _C.check(_C.lib().list_sdf_grid(C.byref(cs), C.byref(wsn), res, bb_min, bb_max, begin, count, out.data_ptr(),
                                        float(sdf_scale), chunk_rows, ws.data_ptr(), ws.numel(), _stream()),
                 "list_sdf_grid")
    return out


def grid_sdf_late(ctx: HotPathContext, weights: KernelWeights, res: int, begin: int, count: int, sdf_scale: float,
                  chunk_rows: int, out: torch.Tensor, workspace: torch.Tensor, late_event: Optional[torch.cuda.Event],
                  late_vols: Sequence[Optional[torch.Tensor]], out_host: Optional[torch.Tensor] = None,
                  bb_min: float = -0.5, bb_max: float = 0.5) -> torch.Tensor:
    """grid_sdf for a context whose fine levels are still being produced on another stream (list_sdf_grid_late):
    late_vols[l] is the reference-layout fp32 device tensor of a level prepare_context skipped (None for the others),
    late_event was recorded after its producer's last write.  out_host (pinned) receives the values chunk by chunk."""
    dev = _require_cuda(ctx.maps_cl, weights.w0)
    if count == 0:
        return out
    if out_host is not None and (not out_host.is_pinned() or out_host.numel() != out.numel() or out_host.dtype != torch.float32):
        raise ValueError("out_host must be a pinned fp32 tensor of B*count elements")
    for l, v in enumerate(late_vols):
        if v is not None and (not v.is_cuda or v.dtype != torch.float32 or not v.is_contiguous()
                              or tuple(v.shape[2:]) != tuple(ctx.vols_cl[l].shape[1:4])):
            raise ValueError(f"late volume {l}: expected a contiguous fp32 NCDHW device tensor matching the context")
    cs, wsn = ctx.struct(), weights.struct()
    raw = _C.ptr_array([0 if v is None else v.data_ptr() for v in late_vols])
    ev = None if late_event is None else late_event.cuda_event
    with torch.cuda.device(dev):
        _C.check(_C.lib().list_sdf_grid_late(C.byref(cs), C.byref(wsn), res, bb_min, bb_max, begin, count, out.data_ptr(),
                                             float(sdf_scale), max(1, min(chunk_rows, count)), workspace.data_ptr(),
                                             workspace.numel(), _stream(), ev, raw,
                                             None if out_host is None else out_host.data_ptr()), "list_sdf_grid_late")
    return out


class HostGridRunner:
    """End-to-end path with HOST buffers (list_sdf_grid_host): pinned reference-layout per-image
    tensors -> H2D -> prep -> grid evaluation -> D2H of the SDF grid."""

    def __init__(self, maps_host, vols_host, trans_host, weights: KernelWeights, res: int, begin: int, count: int,
                 dtype="bf16", chunk_rows: int = DEFAULT_CHUNK, map_size: int = MAP_SIZE):
        self.code = dtype_code(dtype)
        self.maps, self.vols, self.T = maps_host, vols_host, trans_host
        for t in (*maps_host, *vols_host, trans_host):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("host inputs must be contiguous fp32 CPU tensors")
        self.weights, self.res, self.begin, self.count = weights, res, begin, count
        self.chunk_rows = max(1, min(chunk_rows, count))
        self.map_size = map_size
        self.B = maps_host[0].shape[0]
        self.map_ch = _C.i32_array([m.shape[1] for m in maps_host])
        self.map_in = _C.i32_array([m.shape[2] for m in maps_host])
        self.vol_ch = _C.i32_array([v.shape[1] for v in vols_host])
        self.vol_res = _C.i32_array([v.shape[2] for v in vols_host])
        dev = _require_cuda(weights.w0)
        self.dev = dev
        need = _C.lib().list_sdf_grid_host_bytes(self.map_ch, self.map_in, len(maps_host), map_size, len(vols_host),
                                                 self.vol_ch, self.vol_res, self.B, self.code, count, self.chunk_rows)
        self.scratch = torch.empty(need, device=dev, dtype=torch.uint8)
        self.out = torch.empty(self.B, count, dtype=torch.float32).pin_memory()
        self.h2d_bytes = sum(t.numel() * 4 for t in (*maps_host, *vols_host, trans_host))
        self.d2h_bytes = self.out.numel() * 4

    def run(self, sdf_scale: float = 1.0) -> torch.Tensor:
        ws = self.weights.struct()
        with torch.cuda.device(self.dev):
            _C.check(_C.lib().list_sdf_grid_host(
                _C.ptr_array([m.data_ptr() for m in self.maps]), self.map_ch, self.map_in, len(self.maps), self.map_size,
                _C.ptr_array([v.data_ptr() for v in self.vols]), len(self.vols), self.vol_ch, self.vol_res,
                self.T.data_ptr(), self.B, self.code, C.byref(ws), self.res, -0.5, 0.5, self.begin, self.count,
                float(sdf_scale), self.chunk_rows, self.out.data_ptr(), self.scratch.data_ptr(), self.scratch.numel(),
                _stream()), "list_sdf_grid_host")
        return self.out


# ------------------------------------------------------------------------------ mesh extraction
def marching_cubes(grid: torch.Tensor, iso: float = 0.0, negate: bool = True):
    """GPU marching cubes of a dense (res, res, res) fp32 grid on the device (csrc/mcubes.cu).  With negate=True it
    contours -grid, which is what the reference hands to PyMCubes (utils.py:173).  Returns (vertices (nv, 3) fp32 in
    index coordinates, triangles (nt, 3) int32), both on the device; one host read of the two counts in between."""
    dev = _require_cuda(grid)
    if grid.dim() != 3 or not (grid.shape[0] == grid.shape[1] == grid.shape[2]) or grid.dtype != torch.float32:
        raise ValueError("marching_cubes expects a (res, res, res) float32 tensor")
    g = grid.contiguous()
    res = g.shape[0]
    lib = _C.lib()
    need = lib.list_mc_workspace_bytes(res)
    if need == 0:
        raise ValueError(f"marching_cubes: res {res} outside [2, 1024]")
    ws = torch.empty(need, device=dev, dtype=torch.uint8)
    counts = torch.zeros(2, device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        _C.check(lib.list_mc_count(g.data_ptr(), res, float(iso), int(negate), ws.data_ptr(), need, counts.data_ptr(),
                                   _stream()), "list_mc_count")
        nv, nt = (int(x) for x in counts.cpu())
        verts = torch.empty(nv, 3, device=dev, dtype=torch.float32)
        tris = torch.empty(nt, 3, device=dev, dtype=torch.int32)
        _C.check(lib.list_mc_generate(g.data_ptr(), res, float(iso), int(negate), ws.data_ptr(), need, verts.data_ptr(), nv,
                                      tris.data_ptr(), nt, _stream()), "list_mc_generate")
    return verts, tris


# ------------------------------------------------------------------------------ training (fwd + bwd)
class _SdfFunction(torch.autograd.Function):
    """Rows a-2..a-6 + a-9 as one autograd node (fp32).  Inputs are the channels-last per-image
    tensors (built with differentiable torch ops by the caller so that autograd continues into
    the encoders) and the master MLP parameters."""

    @staticmethod
    def forward(ctx, points, raw, trans_mat, maps_cl, w0, b0, w1, b1, w2, b2, w3, b3, *vols_cl):
        dev = _require_cuda(points, trans_mat, maps_cl, w0, *vols_cl)
        hp = HotPathContext(maps_cl.detach().contiguous(), [v.detach().contiguous() for v in vols_cl],
                            trans_mat.detach().to(torch.float32).contiguous(), _C.F32)
        lay = hp.layout
        state = {"fc.fc_0.weight": w0, "fc.fc_0.bias": b0, "fc.fc_1.weight": w1, "fc.fc_1.bias": b1,
                 "fc.fc_2.weight": w2, "fc.fc_2.bias": b2, "fc.fc_out.weight": w3, "fc.fc_out.bias": b3}
        kw = prepare_weights(state, lay, "fp32")
        pts = points.detach().to(torch.float32).contiguous()
        B, N, _ = pts.shape
        X = gather_features(hp, pts, raw)
        sdf, ws = mlp(kw, X, 1.0, return_workspace=True, train=True)
        ctx.hp, ctx.kw, ctx.lay, ctx.raw = hp, kw, lay, raw
        ctx.save_for_backward(pts, X, ws)
        ctx.shapes = (w0.shape, w1.shape, w2.shape, w3.shape)
        ctx.dev = dev
        return sdf.view(B, N)

    @staticmethod
    def backward(ctx, d_sdf):
        pts, X, fwd_ws = ctx.saved_tensors
        hp, kw, lay, dev = ctx.hp, ctx.kw, ctx.lay, ctx.dev
        B, N, _ = pts.shape
        rows = B * N
        lib = _C.lib()
        need = ctx.needs_input_grad
        g = _C.ListGrads()
        d_T = torch.zeros_like(hp.trans_mat) if need[2] else None
        d_maps = torch.zeros_like(hp.maps_cl) if need[3] else None
        d_vols = [torch.zeros_like(v) if need[12 + i] else None for i, v in enumerate(hp.vols_cl)]
        d_w0 = torch.zeros_like(kw.w0)
        d_w1, d_w2, d_w3 = torch.zeros_like(kw.w1), torch.zeros_like(kw.w2), torch.zeros_like(kw.w3)
        d_b0, d_b1, d_b2, d_b3 = (torch.zeros_like(b) for b in (kw.b0, kw.b1, kw.b2, kw.b3))
        g.d_maps = d_maps.data_ptr() if d_maps is not None else None
        for i, dv in enumerate(d_vols):
            g.d_vols[i] = dv.data_ptr() if dv is not None else None
        g.d_trans_mat = d_T.data_ptr() if d_T is not None else None
        g.d_w0, g.d_w1, g.d_w2, g.d_w3 = d_w0.data_ptr(), d_w1.data_ptr(), d_w2.data_ptr(), d_w3.data_ptr()
        g.d_b0, g.d_b1, g.d_b2, g.d_b3 = d_b0.data_ptr(), d_b1.data_ptr(), d_b2.data_ptr(), d_b3.data_ptr()
        cs, wsn = hp.struct(), kw.struct()
        ws = torch.empty(lib.list_bwd_workspace_bytes(C.byref(wsn), rows), device=dev, dtype=torch.uint8)
        dsd = d_sdf.detach().to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            _C.check(lib.list_sdf_bwd(C.byref(cs), C.byref(wsn), pts.data_ptr(), int(ctx.raw), B, N, X.data_ptr(),
                                      X.stride(0), fwd_ws.data_ptr(), dsd.data_ptr(), C.byref(g), ws.data_ptr(),
                                      ws.numel(), _stream()), "list_sdf_bwd")
        perm = torch.from_numpy(lay.perm.astype(np.int64)).to(dev)
        s0, s1, s2, s3 = ctx.shapes
        g_w0 = torch.empty(s0[0], lay.k_out, device=dev, dtype=torch.float32)
        g_w0[:, perm] = d_w0[:, :lay.k_out]
        return (None, None, d_T, d_maps, g_w0.view(s0), d_b0, d_w1.view(s1), d_b1, d_w2.view(s2), d_b2,
                d_w3.view(s3), d_b3, *d_vols)


class _PrepVolumeFn(torch.autograd.Function):
    """NCDHW fp32 volume -> channels-last (B,R,R,R,C) through list_prep_volume; the backward is the inverse layout change
    through list_prep_volume_bwd (a contiguous NCDHW gradient: what the voxel encoder's backward, or a leaf's .grad,
    wants -- a permuted view would be copied element by element by whoever consumes it)."""

    @staticmethod
    def forward(ctx, vol):
        dev = _require_cuda(vol)
        v = vol.detach().to(torch.float32).contiguous()
        B, Cc, R = v.shape[0], v.shape[1], v.shape[2]
        out = torch.empty(B, R, R, R, Cc, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _C.check(_C.lib().list_prep_volume(v.data_ptr(), B, Cc, R, out.data_ptr(), _C.F32, _stream()), "list_prep_volume")
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        B, R, Cc = g.shape[0], g.shape[1], g.shape[4]
        out = torch.empty(B, Cc, R, R, R, device=g.device, dtype=torch.float32)
        with torch.cuda.device(g.device):
            _C.check(_C.lib().list_prep_volume_bwd(g.data_ptr(), B, Cc, R, out.data_ptr(), _stream()), "list_prep_volume_bwd")
        return out


class _PrepMapsFn(torch.autograd.Function):
    """The five NCHW maps -> ONE upsampled channels-last (B,S,S,1024) tensor through list_prep_maps (row a-1, reference
    modules.py:25-35); the backward is list_prep_maps_bwd (adjoint of the upsample + layout in one pass over the
    channels-last gradient)."""

    @staticmethod
    def forward(ctx, map_size, *maps):
        dev = _require_cuda(*maps)
        ms = [m.detach().to(torch.float32).contiguous() for m in maps]
        B = ms[0].shape[0]
        cm = sum(m.shape[1] for m in ms)
        out = torch.empty(B, map_size, map_size, cm, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _C.check(_C.lib().list_prep_maps(_C.ptr_array([m.data_ptr() for m in ms]), _C.i32_array([m.shape[1] for m in ms]),
                                             _C.i32_array([m.shape[2] for m in ms]), len(ms), B, map_size, out.data_ptr(),
                                             _C.F32, _stream()), "list_prep_maps")
        ctx.shapes = [tuple(m.shape) for m in ms]
        ctx.map_size = map_size
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        grads = [torch.empty(shp, device=g.device, dtype=torch.float32) for shp in ctx.shapes]
        with torch.cuda.device(g.device):
            _C.check(_C.lib().list_prep_maps_bwd(g.data_ptr(), _C.i32_array([s[1] for s in ctx.shapes]),
                                                 _C.i32_array([s[2] for s in ctx.shapes]), len(grads), g.shape[0], ctx.map_size,
                                                 _C.ptr_array([t.data_ptr() for t in grads]), _stream()), "list_prep_maps_bwd")
        return (None, *grads)


def prep_maps_autograd(maps: Sequence[torch.Tensor], map_size: int = MAP_SIZE) -> torch.Tensor:
    """Differentiable a-1 + layout: (B,C_i,H_i,W_i) maps -> (B,S,S,sum C) fp32 channels-last."""
    return _PrepMapsFn.apply(map_size, *maps)


def prep_volume_autograd(vol: torch.Tensor) -> torch.Tensor:
    """Differentiable layout change (B,C,R,R,R) -> (B,R,R,R,C) fp32."""
    return _PrepVolumeFn.apply(vol)


def query_sdf_autograd(points, trans_mat, maps_cl, vols_cl, params: dict, raw: bool = True, prefix: str = "fc."):
    """Differentiable a-7: `params` maps the reference's state_dict names to the master tensors."""
    p = params
    return _SdfFunction.apply(points, raw, trans_mat, maps_cl,
                              p[f"{prefix}fc_0.weight"], p[f"{prefix}fc_0.bias"], p[f"{prefix}fc_1.weight"],
                              p[f"{prefix}fc_1.bias"], p[f"{prefix}fc_2.weight"], p[f"{prefix}fc_2.bias"],
                              p[f"{prefix}fc_out.weight"], p[f"{prefix}fc_out.bias"], *vols_cl)


def gemm_f32_tc(A: torch.Tensor, B: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """C (M, N) (+)= A (M, K) . B (N, K)^T with fp32-level accuracy on the tensor cores (list_gemm_f32_tc); fp32, rows
    contiguous."""
    dev = _require_cuda(A, B)
    M, K = A.shape
    N = B.shape[0]
    Cc = out if out is not None else torch.zeros(M, N, device=dev, dtype=torch.float32)
    lo = torch.empty(M * A.stride(0) + N * B.stride(0), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _C.check(_C.lib().list_gemm_f32_tc(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), Cc.data_ptr(), Cc.stride(0),
                                           M, N, K, int(accumulate), lo.data_ptr(), lo.numel() * 4, _stream()), "list_gemm_f32_tc")
    return Cc


def mlp_debug(weights: KernelWeights, X: torch.Tensor, out_div: float = 1.0):
    """Diagnostic: bf16 tensor-core MLP returning (sdf, relu(fc_0), relu(fc_1), relu(fc_2)) in fp32."""
    dev = _require_cuda(X, weights.w0)
    rows = X.shape[0]
    sdf = torch.empty(rows, device=dev, dtype=torch.float32)
    h1 = torch.zeros(rows, weights.w0.shape[0], device=dev, dtype=torch.float32)
    h2 = torch.zeros(rows, weights.w1.shape[0], device=dev, dtype=torch.float32)
    h3 = torch.zeros(rows, weights.w2.shape[0], device=dev, dtype=torch.float32)
    ws = weights.struct()
    with torch.cuda.device(dev):
        _C.check(_C.lib().list_mlp_fwd_debug(C.byref(ws), X.data_ptr(), X.stride(0), rows, sdf.data_ptr(), float(out_div),
                                             h1.data_ptr(), h2.data_ptr(), h3.data_ptr(), _stream()), "list_mlp_fwd_debug")
    return sdf, h1, h2, h3

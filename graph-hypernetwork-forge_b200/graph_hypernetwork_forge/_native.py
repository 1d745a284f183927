"""ctypes binding of libghf_b200.so (C ABI: include/ghf_b200.h).

PyTorch is used for device memory and streams only; every computation on the
HyperGNN forward path happens inside the library's sm_100a kernels.  There is
no fallback: a missing library raises ImportError at first use, a non-CUDA
tensor or a non-Blackwell device raises RuntimeError.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

import torch

PREC_FP32 = 0
PREC_TF32 = 1
PREC_F16 = 2
_PREC = {"fp32": PREC_FP32, "tf32": PREC_TF32, "f16": PREC_F16}
ABI_VERSION = 6

_LIB_PATH = os.environ.get("GHF_LIB") or os.path.join(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib", "libghf_b200.so")
_lib = None
_HOOK_T = ctypes.CFUNCTYPE(c_int, c_void_p)


class ModelDesc(ctypes.Structure):
    _fields_ = [("text_dim", c_int32), ("node_feat_dim", c_int32), ("hidden_dim", c_int32),
                ("num_layers", c_int32), ("char_emb_dim", c_int32), ("gen_hidden", c_int32),
                ("gen_depth", c_int32), ("precision", c_int32), ("ln_eps", c_float)]


def lib_path() -> str:
    return _LIB_PATH


def lib():
    """Load the shared library once and declare every prototype of ghf_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  The B200 path has no CPU or eager-PyTorch fallback.")
    L = ctypes.CDLL(_LIB_PATH)
    P = c_void_p
    sig = {
        "ghf_abi_version": (c_int, []),
        "ghf_last_error": (c_char_p, []),
        "ghf_device_ok": (c_int, []),
        "ghf_set_presync_hook": (c_int, [_HOOK_T, P]),
        "ghf_dedup_texts": (c_int, [P, P, c_int64, P, c_int64, P, P, POINTER(c_int64), P]),
        "ghf_select_edges": (c_int, [P, c_int64, c_int64, c_int64, P, POINTER(c_int64), P]),
        "ghf_text_encode": (c_int, [P, P, P, c_int64, P, c_int, P, P, c_int, P, P]),
        "ghf_linear": (c_int, [P, c_int64, c_int, P, P, c_int, c_int, P, P, P]),
        "ghf_linear_f16out": (c_int, [P, c_int64, c_int, P, P, c_int, c_int, P, P, P, P, P]),
        "ghf_linear_backward": (c_int, [P, c_int64, c_int, P, c_int, c_int, P, P, P, P, P, P, P, P]),
        "ghf_graph_build": (c_int, [P, c_int64, P, c_int64, P, c_int64, c_int32, c_int32, c_int64, c_int64,
                                    c_int32, c_int32, POINTER(c_void_p), P]),
        "ghf_graph_free": (None, [P]),
        "ghf_graph_info": (c_int, [P, POINTER(c_int64)]),
        "ghf_graph_export": (c_int, [P, P, P, c_int64, P, P, P, P, P, P, P, P]),
        "ghf_mp_workspace_bytes": (c_int64, [P, c_int32, c_int]),
        "ghf_mp_layer": (c_int, [P, P, P, P, P, P, P, c_float, c_int, P, P, P, P]),
        "ghf_mp_layer_f16": (c_int, [P, P, P, P, P, P, P, P, P, c_float, c_int, P, P, P, P, P, P]),
        "ghf_mp_layer_f16_range": (c_int, [P, P, P, P, P, P, P, P, P, c_float, c_int, P, P, P, P, P, c_int32, c_int32, P]),
        "ghf_graph_num_phases": (c_int64, [P]),
        "ghf_mp_layer_f16_push": (c_int, [P, P, P, P, P, P, P, P, P, c_float, c_int, P, P, P, P, P, c_int32, c_int32,
                                          P, c_int64, P, c_int32, c_int32, P]),
        "ghf_mark_rows": (c_int, [P, P, c_int64, c_int64, P, P]),
        "ghf_mp_contract": (c_int, [P, P, P, P, P, P, P, c_int, P, c_int, c_int, P, P]),
        "ghf_mp_epilogue_backward": (c_int, [P, P, P, P, P, c_float, P, P, P, P, P, P]),
        "ghf_dropout_offset_advance": (c_int64, [c_int64]),
        "ghf_mp_layer_dropout": (c_int, [P, P, P, P, P, P, P, P, P, c_float, c_int, c_float, c_uint64, c_uint64,
                                         P, P, P, P, P, P]),
        "ghf_mp_epilogue_backward_dropout": (c_int, [P, P, P, P, P, c_float, c_float, c_uint64, c_uint64,
                                                     P, P, P, P, P, P]),
        "ghf_mp_weight_grad": (c_int, [P, P, P, P, P, P, P, c_int, P, P, P, P, P]),
        "ghf_text_encode_backward": (c_int, [P, P, P, c_int64, P, c_int, P, c_int, P, P, P, P, P, P]),
        "ghf_weight_images_bytes": (c_int64, [c_int64, c_int32]),
        "ghf_weight_images_f16": (c_int, [P, P, c_int64, P, P, P, P, P, P, c_int32, P, P]),
        "ghf_mp_layer_images": (c_int, [P, P, P, P, P, P, P, P, c_float, P, P, P, P, P]),
        "ghf_score_pairs": (c_int, [P, c_int64, c_int, P, P, c_int64, P, P]),
        "ghf_score_pairs_backward": (c_int, [P, c_int64, c_int, P, P, c_int64, P, P, P]),
        "ghf_absmax": (c_int, [P, c_int64, P, P]),
        "ghf_convert_f16": (c_int, [P, c_int64, P, P, c_int, P]),
        "ghf_hypergnn_forward_host": (c_int, [POINTER(ModelDesc), POINTER(c_void_p), c_int64, P, c_int64, P,
                                              c_int64, P, P, P, P]),
        "ghf_hypergnn_forward_device": (c_int, [POINTER(ModelDesc), POINTER(c_void_p), c_int64, P, c_int64, P,
                                                c_int64, P, P, P, P]),
        "ghf_copy_async": (c_int, [P, P, c_int64, P]),
        "ghf_weight_generators_scratch_bytes": (c_int64, [c_int64, c_int32, c_int32, c_int32]),
        "ghf_weight_generators": (c_int, [P, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                          POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), P, c_int32, P]),
        "ghf_launch_count": (c_int64, [c_int]),
        "ghf_profile_enable": (c_int, [c_int]),
        "ghf_profile_read": (c_int, [POINTER(ctypes.c_double), POINTER(c_int64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    if L.ghf_abi_version() != ABI_VERSION:
        raise ImportError(f"{_LIB_PATH}: ABI version {L.ghf_abi_version()} != {ABI_VERSION}")
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "ghf_abi_version", "ghf_last_error", "ghf_device_ok", "ghf_set_presync_hook", "ghf_dedup_texts", "ghf_select_edges", "ghf_text_encode", "ghf_linear",
    "ghf_linear_f16out", "ghf_linear_backward",
    "ghf_graph_build", "ghf_graph_free", "ghf_graph_info", "ghf_graph_export", "ghf_mp_workspace_bytes",
    "ghf_mp_layer", "ghf_mp_layer_f16", "ghf_mp_layer_f16_range", "ghf_graph_num_phases", "ghf_mp_layer_f16_push", "ghf_mark_rows", "ghf_mp_contract", "ghf_mp_epilogue_backward", "ghf_mp_weight_grad",
    "ghf_dropout_offset_advance", "ghf_mp_layer_dropout", "ghf_mp_epilogue_backward_dropout",
    "ghf_text_encode_backward", "ghf_weight_images_bytes", "ghf_weight_images_f16", "ghf_mp_layer_images",
    "ghf_score_pairs", "ghf_score_pairs_backward", "ghf_absmax", "ghf_convert_f16", "ghf_hypergnn_forward_host", "ghf_hypergnn_forward_device", "ghf_copy_async", "ghf_weight_generators_scratch_bytes", "ghf_weight_generators",
    "ghf_launch_count", "ghf_profile_enable", "ghf_profile_read",
)


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().ghf_last_error()
        raise RuntimeError(f"{what} failed: {msg.decode(errors='replace') if msg else rc}")


def precision_code(name: str) -> int:
    try:
        return _PREC[name]
    except KeyError:
        raise ValueError(f"precision must be one of {sorted(_PREC)}, got {name!r}") from None


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """All tensors on one CUDA device; the device must be usable by the library."""
    dev = tensors[0].device
    for t in tensors:
        if t.device.type != "cuda":
            raise RuntimeError(
                "graph_hypernetwork_forge (B200 build) runs on CUDA tensors only - there is no CPU fallback; "
                f"got a tensor on {t.device}.  Move the model and inputs to 'cuda'.")
        if t.device != dev:
            raise RuntimeError(f"tensors on different devices: {t.device} vs {dev}")
    with torch.cuda.device(dev):
        _check(lib().ghf_device_ok(), "ghf_device_ok")
    return dev


def _call_with_presync(call, before_sync):
    """Run `call()` (a wrapper around ghf_select_edges / ghf_dedup_texts / ghf_graph_build) with `before_sync` set as
    the one-shot pre-sync hook (ghf_set_presync_hook): the native side invokes it after its kernels are enqueued and
    right before it waits for a size from the device, so whatever `before_sync` enqueues runs through that round
    trip.  `before_sync` runs exactly once - here, afterwards, if the entry point never reached its wait."""
    if before_sync is None:
        return call()
    state = {"ran": False, "exc": None}

    def hook(_arg):
        state["ran"] = True
        try:
            before_sync()
        except BaseException as e:  # noqa: BLE001 - re-raised below, outside the foreign frame
            state["exc"] = e
            return 1
        return 0

    c_hook = _HOOK_T(hook)
    lib().ghf_set_presync_hook(c_hook, None)
    try:
        result = call()
    except RuntimeError:
        if state["exc"] is not None:
            raise state["exc"]
        raise
    finally:
        lib().ghf_set_presync_hook(_HOOK_T(0), None)
    if not state["ran"]:
        before_sync()
    return result


def _ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def _stream(dev) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError(f"expected a float32 tensor, got {t.dtype}")
    return t.detach().contiguous()


def launch_count(reset: bool = False) -> int:
    return int(lib().ghf_launch_count(1 if reset else 0))


def profile_enable(on: bool) -> None:
    _check(lib().ghf_profile_enable(int(on)), "ghf_profile_enable")


def profile_read():
    """-> ({"contraction_ms", "epilogue_ms", "prep_ms"} summed since the last read, number of layer calls)."""
    ms = (ctypes.c_double * 3)()
    n = c_int64(0)
    _check(lib().ghf_profile_read(ms, ctypes.byref(n)), "ghf_profile_read")
    return {"contraction_ms": ms[0], "epilogue_ms": ms[1], "prep_ms": ms[2]}, int(n.value)


# ----------------------------------------------------------------------------- ops
class DropoutState:
    """What one `F.dropout(t, p)` call on a CUDA tensor of `numel` elements would take from torch's default CUDA
    generator: (p, seed, offset).  `draw` reads the generator and advances it exactly as that call would
    (ghf_dropout_offset_advance), so native dropout and torch's own dropout calls share one random stream."""

    __slots__ = ("p", "seed", "offset")

    def __init__(self, p: float, seed: int, offset: int):
        self.p, self.seed, self.offset = float(p), int(seed), int(offset)

    @staticmethod
    def supported(numel: int) -> bool:
        return numel > 0 and numel % 4 == 0     # torch's vectorised kernel; anything else keeps the torch-op path

    @staticmethod
    def draw(p: float, numel: int, device) -> "DropoutState":
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("native dropout reads the generator state on the host: not capturable")
        device = torch.device(device)
        gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
        seed, offset = gen.initial_seed(), gen.get_offset()
        with torch.cuda.device(device):
            adv = int(lib().ghf_dropout_offset_advance(int(numel)))
        if adv < 0 or offset % 4:
            raise RuntimeError(f"native dropout: unsupported tensor size {numel} or generator offset {offset}")
        gen.set_offset(offset + adv)
        return DropoutState(p, seed & 0xFFFFFFFFFFFFFFFF, offset)


class Shadow:
    """fp16 shadow of a float32 feature matrix: `data` (float16) and `scale` (float32[2] on the device) with
    features = data * scale[0]; scale[0] is an exact power of two chosen on the device, scale[1] = max |features|."""

    def __init__(self, data: torch.Tensor, scale: torch.Tensor = None):
        if data.dtype != torch.float16 or not data.is_contiguous():
            raise RuntimeError("Shadow data must be a contiguous float16 tensor")
        self.data = data
        self.scale = torch.zeros(2, dtype=torch.float32, device=data.device) if scale is None else scale

    def rows(self, lo: int, hi: int) -> "Shadow":
        return Shadow(self.data[lo:hi], self.scale)


def linear(x: torch.Tensor, weight: torch.Tensor, bias, relu: bool = False, log_scale=None, want_f16: bool = False,
           out=None, out_shadow=None):
    """exp(log_scale) * act(x @ weight.T + bias); x [M,K] float32 CUDA.  want_f16: -> (y, Shadow of y).
    `out` / `out_shadow`: write into these (contiguous float32 [M,N] / Shadow of [M,N]) instead of allocating."""
    x, weight = _f32(x), _f32(weight)
    dev = x.device
    M, K = x.shape
    N = weight.shape[0]
    if weight.shape[1] != K:
        raise RuntimeError(f"linear: x is [{M},{K}] but weight is {tuple(weight.shape)}")
    bias = None if bias is None else _f32(bias)
    log_scale = None if log_scale is None else _f32(log_scale)
    y = torch.empty((M, N), dtype=torch.float32, device=dev) if out is None else out
    if y.shape != (M, N) or y.dtype != torch.float32 or not y.is_contiguous():
        raise RuntimeError(f"linear: out must be a contiguous float32 [{M},{N}] tensor")
    y16 = None
    if want_f16:
        y16 = out_shadow if out_shadow is not None else Shadow(torch.empty((M, N), dtype=torch.float16, device=dev))
        if y16.data.shape != (M, N):
            raise RuntimeError(f"linear: out_shadow must shadow a [{M},{N}] matrix")
    with torch.cuda.device(dev):
        _check(lib().ghf_linear_f16out(_ptr(x), M, K, _ptr(weight), _ptr(bias), N, int(relu), _ptr(log_scale),
                                       _ptr(y), _ptr(y16.data) if y16 else None, _ptr(y16.scale) if y16 else None,
                                       _stream(dev)), "ghf_linear_f16out")
    return (y, y16) if want_f16 else y


def linear_backward(x, weight, log_scale, y, g_y, relu: bool, need_x: bool = True, need_w: bool = True,
                    need_b: bool = True, need_ls: bool = False):
    """Gradients of `linear` (ghf_linear_backward): -> (g_x, g_w, g_b, g_log_scale), None where not asked for.
    x [M,K], weight [N,K], y = the forward result [M,N], g_y = dL/dy."""
    x, weight, y, g_y = _f32(x), _f32(weight), _f32(y), _f32(g_y)
    dev = x.device
    M, K = x.shape
    N = weight.shape[0]
    if weight.shape[1] != K or y.shape != (M, N) or g_y.shape != (M, N):
        raise RuntimeError(f"linear_backward: x {tuple(x.shape)}, weight {tuple(weight.shape)}, y {tuple(y.shape)}, "
                           f"g_y {tuple(g_y.shape)} do not fit")
    log_scale = None if log_scale is None else _f32(log_scale)
    g_x = torch.empty((M, K), dtype=torch.float32, device=dev) if need_x else None
    g_w = torch.empty((N, K), dtype=torch.float32, device=dev) if need_w else None
    g_b = torch.empty((N,), dtype=torch.float32, device=dev) if need_b else None
    g_ls = torch.empty((1,), dtype=torch.float32, device=dev) if need_ls and log_scale is not None else None
    with torch.cuda.device(dev):
        _check(lib().ghf_linear_backward(_ptr(x), M, K, _ptr(weight), N, int(relu), _ptr(log_scale), _ptr(y),
                                         _ptr(g_y), _ptr(g_x), _ptr(g_w), _ptr(g_b), _ptr(g_ls), _stream(dev)),
               "ghf_linear_backward")
    return g_x, g_w, g_b, g_ls


def weight_generators(text_emb: torch.Tensor, mlps, log_scales, d_in: int, d_out: int):
    """WG:120-143 for several generators in ONE native call (ghf_weight_generators: the hidden Linears of all their
    MLPs are one grouped launch per depth level).  `mlps`: per generator, the Linears of its three MLPs (W_msg,
    W_self, bias order) as lists of (weight, bias); `log_scales`: per generator three [1] tensors.
    -> per generator {"W_msg" [U,d_in,d_out], "W_self", "bias" [U,d_out]}."""
    text_emb = _f32(text_emb)
    dev = text_emb.device
    U, T = text_emb.shape
    n_gen = len(mlps)
    depth = len(mlps[0][0]) - 1
    H = mlps[0][0][0][0].shape[0] if depth > 0 else 0
    flat, keep = [], []
    for gen in mlps:
        if len(gen) != 3 or any(len(m) != depth + 1 for m in gen):
            raise RuntimeError("weight_generators: every generator needs three MLPs of equal depth")
        for m in gen:
            for i, (w, b) in enumerate(m):
                w, b = _f32(w), _f32(b)
                want = (H, T if i == 0 else H) if i < depth else None
                if want is not None and tuple(w.shape) != want:
                    raise RuntimeError(f"weight_generators: hidden Linear {i} is {tuple(w.shape)}, expected {want}")
                keep += [w, b]
                flat += [w.data_ptr(), b.data_ptr()]
    ls = [_f32(t) for g in log_scales for t in g]
    outs = [{"W_msg": torch.empty((U, d_in, d_out), dtype=torch.float32, device=dev),
             "W_self": torch.empty((U, d_in, d_out), dtype=torch.float32, device=dev),
             "bias": torch.empty((U, d_out), dtype=torch.float32, device=dev)} for _ in range(n_gen)]
    n_scr = int(lib().ghf_weight_generators_scratch_bytes(U, H, depth, n_gen))
    scratch = torch.empty(n_scr, dtype=torch.uint8, device=dev)
    params = (c_void_p * len(flat))(*flat)
    lsp = (c_void_p * len(ls))(*[t.data_ptr() for t in ls])
    outp = (c_void_p * (3 * n_gen))(*[o[k].data_ptr() for o in outs for k in ("W_msg", "W_self", "bias")])
    with torch.cuda.device(dev):
        _check(lib().ghf_weight_generators(_ptr(text_emb), U, T, H, depth, n_gen, d_in, d_out, params, lsp, outp,
                                           _ptr(scratch), 0, _stream(dev)), "ghf_weight_generators")
    return outs


def mark_rows(ids: torch.Tensor, num_rows: int, subset=None) -> torch.Tensor:
    """uint8 [num_rows]: 1 where the row id occurs in `ids` (int64; only the entries listed in `subset` when given):
    the rows of the node table a rank gathers from."""
    dev = ids.device
    if ids.dtype != torch.int64 or not ids.is_contiguous():
        raise RuntimeError("mark_rows: ids must be a contiguous int64 tensor")
    mask = torch.empty(max(num_rows, 1), dtype=torch.uint8, device=dev)
    n = ids.numel() if subset is None else subset.numel()
    with torch.cuda.device(dev):
        _check(lib().ghf_mark_rows(_ptr(ids), _ptr(subset), n, num_rows, _ptr(mask), _stream(dev)), "ghf_mark_rows")
    return mask[:num_rows]


class PeerPush:
    """Arguments of the epilogue's stores into the peers' tables (`Graph.mp_layer(push=...)`): `mask` uint8
    [world, local rows] (which peer reads which of my rows), `tables` int64 [world] device array of table base
    pointers, this rank's index."""

    def __init__(self, mask: torch.Tensor, tables: torch.Tensor, rank: int):
        if mask.dtype != torch.uint8 or mask.dim() != 2 or not mask.is_contiguous():
            raise RuntimeError("PeerPush: mask must be a contiguous uint8 [world, local rows] tensor")
        if tables.dtype != torch.int64 or tables.numel() != mask.shape[0]:
            raise RuntimeError("PeerPush: tables must hold one int64 pointer per rank")
        self.mask, self.tables, self.rank, self.world = mask, tables, int(rank), int(mask.shape[0])


def copy_async(dst: torch.Tensor, src: torch.Tensor) -> None:
    """Stream-ordered raw copy src -> dst (same byte size, contiguous); dst may be a peer-mapped view of another
    GPU's buffer (torch symmetric memory): the copy engines move the rows over NVLink, no SM is used."""
    if not (dst.is_contiguous() and src.is_contiguous()) or dst.numel() * dst.element_size() != src.numel() * src.element_size():
        raise RuntimeError("copy_async: contiguous tensors of equal byte size")
    dev = src.device
    with torch.cuda.device(dev):
        _check(lib().ghf_copy_async(_ptr(dst), _ptr(src), src.numel() * src.element_size(), _stream(dev)),
               "ghf_copy_async")


def absmax(x: torch.Tensor, shadow: Shadow) -> None:
    """shadow.scale[1] = max |x| (float32 CUDA, numel % 8 == 0)."""
    x = _f32(x)
    with torch.cuda.device(x.device):
        _check(lib().ghf_absmax(_ptr(x), x.numel(), _ptr(shadow.scale), _stream(x.device)), "ghf_absmax")


def to_f16(x: torch.Tensor, shadow: Shadow, have_amax: bool = False) -> Shadow:
    """shadow.data = fp16(x * 2^k) with the scale picked on the device from max |x| (from shadow.scale[1] when
    `have_amax`, e.g. after an all-reduce over ranks; computed here otherwise)."""
    x = _f32(x)
    if shadow.data.shape != x.shape:
        raise RuntimeError("to_f16: shadow and x must have the same shape")
    with torch.cuda.device(x.device):
        _check(lib().ghf_convert_f16(_ptr(x), x.numel(), _ptr(shadow.data), _ptr(shadow.scale), int(have_amax),
                                     _stream(x.device)), "ghf_convert_f16")
    return shadow


def dedup_texts(utf8: torch.Tensor, offsets: torch.Tensor, subset=None, before_sync=None):
    """-> (rel_ids int32, first_edge int64 [U]) for packed strings on the device.  rel_ids has one entry per
    string, or per entry of `subset` (ascending string ids, int32 storage read as uint32) when given.
    `before_sync`: see `_call_with_presync`."""
    dev = utf8.device
    E = offsets.numel() - 1
    n = E if subset is None else subset.numel()
    if subset is not None and subset.dtype != torch.int32:
        raise RuntimeError("subset must be an int32 tensor")
    rel = torch.empty(n, dtype=torch.int32, device=dev)
    first = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    nu = c_int64(0)
    with torch.cuda.device(dev):
        _call_with_presync(lambda: _check(
            lib().ghf_dedup_texts(_ptr(utf8), _ptr(offsets), E, _ptr(subset), n if subset is not None else 0,
                                  _ptr(rel), _ptr(first), ctypes.byref(nu), _stream(dev)), "ghf_dedup_texts"), before_sync)
    return rel, first[: nu.value]


def select_edges(edge_index: torch.Tensor, dst_lo: int, dst_hi: int, before_sync=None) -> torch.Tensor:
    """Ascending ids (int32 storage, uint32 values) of the edges whose destination lies in [dst_lo, dst_hi).
    `before_sync`: see `_call_with_presync`."""
    dev = edge_index.device
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise RuntimeError("edge_index must be an int64 tensor of shape [2, E]")
    edge_index = edge_index.contiguous()
    E = edge_index.shape[1]
    ids = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    n = c_int64(0)
    with torch.cuda.device(dev):
        _call_with_presync(lambda: _check(
            lib().ghf_select_edges(_ptr(edge_index), E, int(dst_lo), int(dst_hi), _ptr(ids), ctypes.byref(n),
                                   _stream(dev)), "ghf_select_edges"), before_sync)
    return ids[: n.value]


def text_encode(utf8, offsets, index, num, char_emb, proj_w, proj_b) -> torch.Tensor:
    dev = utf8.device
    char_emb, proj_w, proj_b = _f32(char_emb), _f32(proj_w), _f32(proj_b)
    C, T = char_emb.shape[1], proj_w.shape[0]
    out = torch.empty((num, T), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _check(lib().ghf_text_encode(_ptr(utf8), _ptr(offsets), _ptr(index), num, _ptr(char_emb), C, _ptr(proj_w),
                                     _ptr(proj_b), T, _ptr(out), _stream(dev)), "ghf_text_encode")
    return out


def text_encode_backward(utf8, offsets, index, num, char_emb, proj_w, out, g_out):
    """Parameter gradients of `text_encode` -> (g_char_emb [128,C], g_proj_w [T,C], g_proj_b [T])."""
    dev = utf8.device
    char_emb, proj_w, out, g_out = _f32(char_emb), _f32(proj_w), _f32(out), _f32(g_out)
    C, T = char_emb.shape[1], proj_w.shape[0]
    g_emb, g_w = torch.empty_like(char_emb), torch.empty_like(proj_w)
    g_b = torch.empty(T, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _check(lib().ghf_text_encode_backward(_ptr(utf8), _ptr(offsets), _ptr(index), num, _ptr(char_emb), C,
                                              _ptr(proj_w), T, _ptr(out), _ptr(g_out), _ptr(g_emb), _ptr(g_w),
                                              _ptr(g_b), _stream(dev)), "ghf_text_encode_backward")
    return g_emb, g_w, g_b


def _pair_ids(emb, heads, tails):
    if heads.dtype != torch.int64 or tails.dtype != torch.int64 or heads.shape != tails.shape or heads.dim() != 1:
        raise RuntimeError("heads and tails must be 1-D int64 tensors of the same length")
    return heads.contiguous(), tails.contiguous()


def score_pairs(emb: torch.Tensor, heads: torch.Tensor, tails: torch.Tensor) -> torch.Tensor:
    """out[b] = <emb[heads[b]], emb[tails[b]]> without materialising the gathered rows."""
    emb = _f32(emb)
    heads, tails = _pair_ids(emb, heads, tails)
    dev = emb.device
    out = torch.empty(heads.numel(), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _check(lib().ghf_score_pairs(_ptr(emb), emb.shape[0], emb.shape[1], _ptr(heads), _ptr(tails), heads.numel(),
                                     _ptr(out), _stream(dev)), "ghf_score_pairs")
    return out


def score_pairs_backward(emb, heads, tails, g_out) -> torch.Tensor:
    emb, g_out = _f32(emb), _f32(g_out)
    heads, tails = _pair_ids(emb, heads, tails)
    dev = emb.device
    g_emb = torch.empty_like(emb)
    with torch.cuda.device(dev):
        _check(lib().ghf_score_pairs_backward(_ptr(emb), emb.shape[0], emb.shape[1], _ptr(heads), _ptr(tails),
                                              heads.numel(), _ptr(g_out), _ptr(g_emb), _stream(dev)),
               "ghf_score_pairs_backward")
    return g_emb


def weight_images(Zm, Zs, W3m, b3m, ls_m, W3s, b3s, ls_s, hidden_dim: int) -> torch.Tensor:
    """ghf_weight_images_f16: the last Linear of the W_msg / W_self generators written as the fp16 operand images of
    the hidden-64/256 contraction (inputs [R,128]; -> uint8 tensor for `Graph.mp_layer_images`)."""
    Zm, Zs, W3m, b3m, W3s, b3s, ls_m, ls_s = map(_f32, (Zm, Zs, W3m, b3m, W3s, b3s, ls_m, ls_s))
    dev, R = Zm.device, Zm.shape[0]
    if Zm.shape != (R, 128) or Zs.shape != (R, 128) or W3m.shape != (hidden_dim ** 2, 128) or W3s.shape != W3m.shape:
        raise RuntimeError("weight_images: inputs must be [R,128], parameters [d*d,128]")
    n = int(lib().ghf_weight_images_bytes(R, hidden_dim))
    if n < 0:
        raise RuntimeError(f"weight_images: hidden_dim {hidden_dim} has no streamed-weights engine")
    buf = torch.empty(n + 1024, dtype=torch.uint8, device=dev)
    off = (-buf.data_ptr()) % 1024
    images = buf[off:off + n]
    with torch.cuda.device(dev):
        _check(lib().ghf_weight_images_f16(_ptr(Zm), _ptr(Zs), R, _ptr(W3m), _ptr(b3m), _ptr(ls_m), _ptr(W3s),
                                           _ptr(b3s), _ptr(ls_s), hidden_dim, _ptr(images), _stream(dev)),
               "ghf_weight_images_f16")
    return images


class Graph:
    """Owner of a ghf_graph handle (in-degree, dst-CSR, relation-grouped edge order)."""

    def __init__(self, edge_index: torch.Tensor, rel_ids: torch.Tensor, num_nodes: int, num_rel: int,
                 hidden_dim: int, dst_lo: int = 0, dst_hi=None, sb_nodes: int = 0, unit_edges: int = 0,
                 edge_ids=None, before_sync=None):
        """`edge_ids` (from `select_edges`): build from those edges only; rel_ids is then indexed like edge_ids.
        `before_sync`: see `_call_with_presync`."""
        dev = edge_index.device
        if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise RuntimeError("edge_index must be an int64 tensor of shape [2, E]")
        if rel_ids.dtype != torch.int32:
            raise RuntimeError("rel_ids must be int32")
        edge_index = edge_index.contiguous()
        self.device = dev
        self.num_nodes, self.hidden_dim, self.num_rel = int(num_nodes), int(hidden_dim), int(num_rel)
        self.dst_lo = int(dst_lo)
        self.dst_hi = int(num_nodes if dst_hi is None else dst_hi)
        self._h = c_void_p()
        E = edge_index.shape[1]
        n_sub = 0 if edge_ids is None else edge_ids.numel()
        if rel_ids.numel() != (E if edge_ids is None else n_sub):
            raise RuntimeError("rel_ids must have one entry per edge (or per selected edge)")
        rel_c = rel_ids.contiguous()
        with torch.cuda.device(dev):
            _call_with_presync(lambda: _check(
                lib().ghf_graph_build(_ptr(edge_index), E, _ptr(edge_ids), n_sub, _ptr(rel_c), self.num_nodes,
                                      self.num_rel, self.hidden_dim, self.dst_lo, self.dst_hi, int(sb_nodes),
                                      int(unit_edges), ctypes.byref(self._h), _stream(dev)), "ghf_graph_build"),
                before_sync)
        self._inputs = (edge_index, edge_ids, rel_ids.contiguous())   # export() recomputes the permutation from them
        info = (c_int64 * 6)()
        _check(lib().ghf_graph_info(self._h, info), "ghf_graph_info")
        (self.num_kept, self.num_units, self.sb_nodes, self.unit_edges, self.num_local, self.bytes) = map(int, info)
        self.num_phases = int(lib().ghf_graph_num_phases(self._h))
        self._workspace = {}

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.ghf_graph_free(h)

    def export(self):
        """Integer tables for parity checks (device tensors)."""
        dev = self.device
        perm = torch.empty(self.num_kept, dtype=torch.int64, device=dev)
        indeg = torch.empty(self.num_local, dtype=torch.int32, device=dev)
        rowptr = torch.empty(self.num_local + 1, dtype=torch.int64, device=dev)
        us = torch.empty(self.num_units, dtype=torch.int32, device=dev)
        uc = torch.empty_like(us)
        ur = torch.empty_like(us)
        with torch.cuda.device(dev):
            ei, ids, rel = self._inputs
            _check(lib().ghf_graph_export(self._h, _ptr(ei), _ptr(ids), 0 if ids is None else ids.numel(), _ptr(rel),
                                          _ptr(perm), _ptr(indeg), _ptr(rowptr), _ptr(us), _ptr(uc), _ptr(ur),
                                          _stream(dev)), "ghf_graph_export")
        return {"perm": perm, "indeg": indeg, "rowptr": rowptr, "unit_start": us, "unit_count": uc,
                "unit_rel": ur}

    def in_degree(self) -> torch.Tensor:
        """int32 [local nodes] (multi-edges counted), cached."""
        deg = getattr(self, "_indeg", None)
        if deg is None:
            deg = self._indeg = self.export()["indeg"]
        return deg

    def workspace(self, precision: int) -> torch.Tensor:
        key = precision
        ws = self._workspace.get(key)
        if ws is None:
            n = int(lib().ghf_mp_workspace_bytes(self._h, self.hidden_dim, precision))
            ws = torch.empty(n, dtype=torch.uint8, device=self.device)
            self._workspace[key] = ws
        return ws

    def phase_rows(self, phase_lo: int, phase_hi: int):
        """Local row range [r0, r1) of the super-blocks [phase_lo, phase_hi)."""
        return min(phase_lo * self.sb_nodes, self.num_local), min(phase_hi * self.sb_nodes, self.num_local)

    def mp_layer(self, h, W_msg, W_self, bias, ln_w, ln_b, eps: float, precision: int, out=None,
                 want_upd: bool = False, h16=None, out16=None, h_row0=None, phases=None, push=None, dropout=None):
        """One message-passing layer on this graph's destination range -> (out, upd or None).

        `h16` (Shadow of [N, d], optional) is the fp16 shadow of `h` the PREC_F16 contraction gathers from (made
        inside when absent); `out16` (Shadow of [local nodes, d], optional) receives the shadow of `out`.
        `h_row0` (PREC_F16 with `h16` only): `h` holds just the rows [h_row0, h_row0 + len(h)) of the fp32 features
        - enough, because with a shadow the fp32 rows are read only at this graph's own destinations (residual).
        `phases` = (lo, hi): only the super-blocks [lo, hi) of the graph; `out`, `out16`, `upd` still cover all local
        rows, of which `phase_rows(lo, hi)` are written.  `push` (PeerPush, with `out16`): the epilogue kernel also
        stores each fp16 row into the tables of the peers that read it.  `dropout` (DropoutState): training-mode
        dropout between the ReLU and the LayerNorm, inside the row epilogue (whole graph, no `phases` / `push`)."""
        dev = self.device
        h, W_msg, W_self, bias = _f32(h), _f32(W_msg), _f32(W_self), _f32(bias)
        ln_w, ln_b = _f32(ln_w), _f32(ln_b)
        d = self.hidden_dim
        h_ptr = _ptr(h)
        if h_row0 is not None:
            if precision != PREC_F16 or h16 is None:
                raise RuntimeError("h_row0 needs precision f16 and the fp16 shadow h16")
            if h.dim() != 2 or h.shape[1] != d or h_row0 > self.dst_lo or h_row0 + h.shape[0] < self.dst_hi:
                raise RuntimeError(f"h rows [{h_row0},{h_row0 + h.shape[0]}) must cover [{self.dst_lo},{self.dst_hi})")
            h_ptr = c_void_p(h.data_ptr() - int(h_row0) * d * 4)     # row dst_lo of this pointer is h[dst_lo - h_row0]
        elif h.shape != (self.num_nodes, d):
            raise RuntimeError(f"h must be [{self.num_nodes},{d}], got {tuple(h.shape)}")
        if W_msg.shape != (self.num_rel, d, d) or W_self.shape != (self.num_rel, d, d) or \
                bias.shape != (self.num_rel, d):
            raise RuntimeError("relation weights must be [R,d,d], [R,d,d], [R,d]")
        if out is None:
            out = torch.empty((self.num_local, d), dtype=torch.float32, device=dev)
        elif out.shape != (self.num_local, d) or not out.is_contiguous() or out.dtype != torch.float32:
            raise RuntimeError("out must be a contiguous float32 [local nodes, d] tensor")
        upd = torch.empty_like(out) if want_upd else None
        for name, t, rows in (("h16", h16, self.num_nodes), ("out16", out16, self.num_local)):
            if t is not None and (not isinstance(t, Shadow) or t.data.shape != (rows, d)):
                raise RuntimeError(f"{name} must be a Shadow of a [{rows},{d}] matrix")
        ws = self.workspace(precision)
        p_lo, p_hi = (0, self.num_phases) if phases is None else phases
        if push is not None:
            if dropout is not None:
                raise RuntimeError("dropout runs on the whole graph of one GPU")
            if out16 is None or push.mask.shape[1] < self.num_local:
                raise RuntimeError("push needs out16 and a mask row per local node")
            with torch.cuda.device(dev):
                _check(lib().ghf_mp_layer_f16_push(self._h, h_ptr, _ptr(h16.data) if h16 else None,
                                                   _ptr(h16.scale) if h16 else None, _ptr(W_msg), _ptr(W_self), _ptr(bias),
                                                   _ptr(ln_w), _ptr(ln_b), float(eps), precision, _ptr(out),
                                                   _ptr(out16.data), _ptr(out16.scale), _ptr(upd), _ptr(ws), int(p_lo),
                                                   int(p_hi), _ptr(push.mask), int(push.mask.shape[1]),
                                                   _ptr(push.tables), push.world, push.rank, _stream(dev)),
                       "ghf_mp_layer_f16_push")
            return out, upd
        if dropout is not None:
            if phases is not None or h_row0 is not None:
                raise RuntimeError("dropout runs on the whole graph of one GPU")
            with torch.cuda.device(dev):
                _check(lib().ghf_mp_layer_dropout(self._h, h_ptr, _ptr(h16.data) if h16 else None,
                                                  _ptr(h16.scale) if h16 else None, _ptr(W_msg), _ptr(W_self),
                                                  _ptr(bias), _ptr(ln_w), _ptr(ln_b), float(eps), precision,
                                                  dropout.p, dropout.seed, dropout.offset, _ptr(out),
                                                  _ptr(out16.data) if out16 else None,
                                                  _ptr(out16.scale) if out16 else None, _ptr(upd), _ptr(ws),
                                                  _stream(dev)), "ghf_mp_layer_dropout")
            return out, upd
        with torch.cuda.device(dev):
            _check(lib().ghf_mp_layer_f16_range(self._h, h_ptr, _ptr(h16.data) if h16 else None,
                                                _ptr(h16.scale) if h16 else None, _ptr(W_msg), _ptr(W_self), _ptr(bias),
                                                _ptr(ln_w), _ptr(ln_b), float(eps), precision, _ptr(out),
                                                _ptr(out16.data) if out16 else None,
                                                _ptr(out16.scale) if out16 else None, _ptr(upd), _ptr(ws), int(p_lo),
                                                int(p_hi), _stream(dev)), "ghf_mp_layer_f16_range")
        return out, upd

    def mp_layer_images(self, h, images, bias, ln_w, ln_b, eps: float, h16=None, out16=None) -> torch.Tensor:
        """`mp_layer` (PREC_F16, hidden 64 / 256) on operand images made by `weight_images`."""
        dev, d = self.device, self.hidden_dim
        h, bias, ln_w, ln_b = _f32(h), _f32(bias), _f32(ln_w), _f32(ln_b)
        if h.shape != (self.num_nodes, d) or bias.shape != (self.num_rel, d):
            raise RuntimeError("mp_layer_images: h must be [N,d], bias [R,d]")
        out = torch.empty((self.num_local, d), dtype=torch.float32, device=dev)
        ws = self.workspace(PREC_F16)
        with torch.cuda.device(dev):
            _check(lib().ghf_mp_layer_images(self._h, _ptr(h), _ptr(h16.data) if h16 else None,
                                             _ptr(h16.scale) if h16 else None, _ptr(images), _ptr(bias), _ptr(ln_w),
                                             _ptr(ln_b), float(eps), _ptr(out), _ptr(out16.data) if out16 else None,
                                             _ptr(out16.scale) if out16 else None, _ptr(ws), _stream(dev)),
                   "ghf_mp_layer_images")
        return out

    # ---- gradients -------------------------------------------------------------------
    def reversed(self) -> "Graph":
        """The graph of the reversed edge list (same relations), built on first use: the contraction over it
        carries gradients from destinations back to sources."""
        rev = getattr(self, "_reversed", None)
        if rev is None:
            ei, ids, rel = self._inputs
            if ids is not None or self.dst_lo != 0 or self.dst_hi != self.num_nodes:
                raise RuntimeError("gradients need a full-range graph (one GPU)")
            rev = Graph(ei.flip(0).contiguous(), rel, self.num_nodes, self.num_rel, self.hidden_dim)
            self._reversed = rev
        return rev

    def contract(self, x, W_msg, W_self, bias, precision: int, x16=None, out=None, accumulate: bool = False,
                 transposed: bool = False):
        """Raw per-destination sums of x_u W_msg[r] + x_v W_self[r] + bias[r] -> [local nodes, d]; with
        `accumulate` they are added to `out`.  PREC_F16 at hidden 128 only: `transposed` uses W[r]^T, and W_msg,
        W_self or bias may be None (zeros; the rows of an absent half are not even gathered)."""
        dev, d = self.device, self.hidden_dim
        if x is None:                                        # PREC_F16 with a shadow: the fp32 rows are never read
            if x16 is None or precision != PREC_F16:
                raise RuntimeError("contract: x may be omitted only with precision f16 and its fp16 shadow x16")
            x_ptr = _ptr(x16.data)                           # any valid pointer
        else:
            x = _f32(x)
            if x.shape != (self.num_nodes, d):
                raise RuntimeError(f"x must be [{self.num_nodes},{d}], got {tuple(x.shape)}")
            x_ptr = _ptr(x)
        if x16 is not None and (not isinstance(x16, Shadow) or x16.data.shape != (self.num_nodes, d)):
            raise RuntimeError(f"x16 must be a Shadow of a [{self.num_nodes},{d}] matrix")
        W_msg, W_self, bias = (None if t is None else _f32(t) for t in (W_msg, W_self, bias))
        for t, shape in ((W_msg, (self.num_rel, d, d)), (W_self, (self.num_rel, d, d)), (bias, (self.num_rel, d))):
            if t is not None and t.shape != shape:
                raise RuntimeError("relation weights must be [R,d,d], [R,d,d], [R,d]")
        if out is None:
            if accumulate:
                raise RuntimeError("contract: accumulate needs `out`")
            out = torch.empty((self.num_local, d), dtype=torch.float32, device=dev)
        elif out.shape != (self.num_local, d) or out.dtype != torch.float32 or not out.is_contiguous():
            raise RuntimeError("contract: out must be a contiguous float32 [local nodes, d] tensor")
        ws = self.workspace(precision)
        with torch.cuda.device(dev):
            _check(lib().ghf_mp_contract(self._h, x_ptr, _ptr(x16.data) if x16 else None,
                                         _ptr(x16.scale) if x16 else None, _ptr(W_msg), _ptr(W_self), _ptr(bias),
                                         precision, _ptr(out), int(accumulate), int(transposed), _ptr(ws),
                                         _stream(dev)), "ghf_mp_contract")
        return out

    def epilogue_backward(self, g_out, upd, h, ln_w, eps: float, want_shadow: bool = False, dropout=None,
                          h_row0=None, shadow_out=None):
        """-> (g_pre, g_acc, g_ln_w, g_ln_b, Shadow of g_acc or None); see ghf_mp_epilogue_backward.  `dropout`: the
        DropoutState of the forward call (the mask is regenerated from it).  `h_row0`: `h` holds only the rows
        [h_row0, h_row0 + len(h)) of the features (they are read at this graph's own destinations only).
        `shadow_out` (a Shadow of [local nodes, d], e.g. the local rows of a full table): receives the shadow of g_acc."""
        dev, d = self.device, self.hidden_dim
        g_out, upd, h, ln_w = _f32(g_out), _f32(upd), _f32(h), _f32(ln_w)
        if g_out.shape != (self.num_local, d) or upd.shape != g_out.shape:
            raise RuntimeError("epilogue_backward: shape mismatch")
        h_ptr = _ptr(h)
        if h_row0 is not None:
            if h.dim() != 2 or h.shape[1] != d or h_row0 > self.dst_lo or h_row0 + h.shape[0] < self.dst_hi:
                raise RuntimeError(f"h rows [{h_row0},{h_row0 + h.shape[0]}) must cover [{self.dst_lo},{self.dst_hi})")
            h_ptr = c_void_p(h.data_ptr() - int(h_row0) * d * 4)
        elif h.shape != (self.num_nodes, d):
            raise RuntimeError("epilogue_backward: shape mismatch")
        g_pre, g_acc = torch.empty_like(g_out), torch.empty_like(g_out)
        g_w, g_b = torch.empty_like(ln_w), torch.empty_like(ln_w)
        if shadow_out is not None:
            if not isinstance(shadow_out, Shadow) or shadow_out.data.shape != g_acc.shape:
                raise RuntimeError("epilogue_backward: shadow_out must shadow a [local nodes, d] matrix")
            g16 = shadow_out
        else:
            g16 = Shadow(torch.empty(g_acc.shape, dtype=torch.float16, device=dev)) if want_shadow else None
        with torch.cuda.device(dev):
            if dropout is not None:
                _check(lib().ghf_mp_epilogue_backward_dropout(
                    self._h, _ptr(g_out), _ptr(upd), h_ptr, _ptr(ln_w), float(eps), dropout.p, dropout.seed,
                    dropout.offset, _ptr(g_pre), _ptr(g_acc), _ptr(g_w), _ptr(g_b), _ptr(g16.scale) if g16 else None,
                    _stream(dev)), "ghf_mp_epilogue_backward_dropout")
            else:
                _check(lib().ghf_mp_epilogue_backward(self._h, _ptr(g_out), _ptr(upd), h_ptr, _ptr(ln_w),
                                                      float(eps), _ptr(g_pre), _ptr(g_acc), _ptr(g_w), _ptr(g_b),
                                                      _ptr(g16.scale) if g16 else None, _stream(dev)),
                       "ghf_mp_epilogue_backward")
        if g16 is not None:
            to_f16(g_acc, g16, have_amax=True)
        return g_pre, g_acc, g_w, g_b, g16

    def weight_grad(self, h, g_acc, precision: int, h16=None, g16=None):
        """-> (g_W_msg [R,d,d], g_W_self [R,d,d], g_bias [R,d]); see ghf_mp_weight_grad."""
        dev, d = self.device, self.hidden_dim
        g_acc = _f32(g_acc)
        if h is None:                                        # PREC_F16 with shadows: the fp32 rows are never read
            if h16 is None or precision != PREC_F16:
                raise RuntimeError("weight_grad: h may be omitted only with precision f16 and its fp16 shadow h16")
            h = h16.data
        else:
            h = _f32(h)
            if h.shape != (self.num_nodes, d):
                raise RuntimeError("weight_grad: shape mismatch")
        if g_acc.shape != (self.num_local, d):
            raise RuntimeError("weight_grad: shape mismatch")
        gm = torch.empty((self.num_rel, d, d), dtype=torch.float32, device=dev)
        gs = torch.empty_like(gm)
        gb = torch.empty((self.num_rel, d), dtype=torch.float32, device=dev)
        ws = self.workspace(precision)
        with torch.cuda.device(dev):
            _check(lib().ghf_mp_weight_grad(self._h, _ptr(h), _ptr(h16.data) if h16 else None,
                                            _ptr(h16.scale) if h16 else None, _ptr(g_acc),
                                            _ptr(g16.data) if g16 else None, _ptr(g16.scale) if g16 else None,
                                            precision, _ptr(gm), _ptr(gs), _ptr(gb), _ptr(ws), _stream(dev)),
                   "ghf_mp_weight_grad")
        return gm, gs, gb

def forward_host(desc: ModelDesc, params, node_features, edge_index, utf8, offsets, out, device):
    """ghf_hypergnn_forward_host: HOST numpy/pinned buffers in, host buffer out (copies inside)."""
    arr = (c_void_p * len(params))(*[p.data_ptr() for p in params])
    N = node_features.shape[0]
    E = edge_index.shape[1]
    with torch.cuda.device(device):
        _check(lib().ghf_hypergnn_forward_host(
            ctypes.byref(desc), arr, len(params), c_void_p(node_features.data_ptr()), N,
            c_void_p(edge_index.data_ptr()), E, c_void_p(utf8.data_ptr()), c_void_p(offsets.data_ptr()),
            c_void_p(out.data_ptr()), _stream(device)), "ghf_hypergnn_forward_host")
    return out


def forward_device(desc: ModelDesc, params, node_features, edge_index, utf8, offsets, out=None):
    """ghf_hypergnn_forward_device: the whole forward on CUDA tensors in ONE native call -> out [N, hidden]."""
    dev = node_features.device
    node_features = _f32(node_features)
    edge_index = edge_index.contiguous()
    if edge_index.dtype != torch.int64 or utf8.dtype != torch.uint8 or offsets.dtype != torch.int64:
        raise RuntimeError("edge_index/offsets must be int64 and utf8 uint8")
    N, E = node_features.shape[0], edge_index.shape[1]
    if out is None:
        out = torch.empty((N, desc.hidden_dim), dtype=torch.float32, device=dev)
    arr = (c_void_p * len(params))(*[p.data_ptr() for p in params])
    with torch.cuda.device(dev):
        _check(lib().ghf_hypergnn_forward_device(
            ctypes.byref(desc), arr, len(params), _ptr(node_features), N, _ptr(edge_index), E, _ptr(utf8.contiguous()),
            _ptr(offsets.contiguous()), _ptr(out), _stream(dev)), "ghf_hypergnn_forward_device")
    return out

"""ctypes view of liblpr381.so, the C++ host layer (product side)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "linear_programming_solver_lpr381_b200")
_lib = None


def lib():
    global _lib
    if _lib is None:
        C.CDLL(os.path.join(PKG, "liblpx.so"), mode=C.RTLD_GLOBAL)
        L = C.CDLL(os.path.join(PKG, "liblpr381.so"))
        L.lpr_solve_text.restype = C.c_void_p
        for f in ("lpr_text_error", "lpr_text_log", "lpr_text_report", "lpr_text_summary"):
            getattr(L, f).restype = C.c_char_p
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("lpr_text_code", "lpr_text_chunks", "lpr_text_highlighted", "lpr_text_free"):
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("lpr_fmt_custom", "lpr_fmt_fixed"):
            getattr(L, f).restype = C.c_char_p
            getattr(L, f).argtypes = [C.c_double, C.c_int]
        L.lpr_fmt_roundtrip.restype = C.c_char_p
        L.lpr_fmt_roundtrip.argtypes = [C.c_double]
        L.lpr_math_round.restype = C.c_double
        L.lpr_math_round.argtypes = [C.c_double, C.c_int]
        L.lpr_normalize_key.restype = C.c_char_p
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.lpr_text_result_dims.argtypes = [C.c_void_p, ip, ip, ip, ip]
        L.lpr_text_tableau.restype = dp
        L.lpr_text_solution.restype = dp
        L.lpr_text_basis.restype = ip
        L.lpr_text_z.restype = C.c_double
        for f in ("lpr_text_tableau", "lpr_text_solution", "lpr_text_basis", "lpr_text_z", "lpr_text_cut_count"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.lpr_text_cut.argtypes = [C.c_void_p, C.c_int, dp, dp]
        _lib = L
    return _lib


def solve_text(text, algorithm):
    L = lib()
    h = L.lpr_solve_text(text.encode("utf-8"), algorithm.encode("utf-8"))
    try:
        out = dict(code=L.lpr_text_code(h), error=L.lpr_text_error(h).decode("utf-8"),
                   log=L.lpr_text_log(h).decode("utf-8"), report=L.lpr_text_report(h).decode("utf-8"),
                   summary=L.lpr_text_summary(h).decode("utf-8"), chunks=L.lpr_text_chunks(h),
                   highlighted=L.lpr_text_highlighted(h))
        rows, cols, nx, nb = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.lpr_text_result_dims(h, C.byref(rows), C.byref(cols), C.byref(nx), C.byref(nb))
        out["tableau"] = (np.ctypeslib.as_array(L.lpr_text_tableau(h), (rows.value, cols.value)).copy()
                          if rows.value else None)
        out["x"] = np.ctypeslib.as_array(L.lpr_text_solution(h), (nx.value,)).copy() if nx.value else None
        out["basis"] = np.ctypeslib.as_array(L.lpr_text_basis(h), (nb.value,)).copy() if nb.value else None
        out["z"] = L.lpr_text_z(h)
        out["masks"], out["chunk_len"] = _masks(L.lpr_text_mask, L.lpr_text_chunk_len, h, out["chunks"])
        cuts = []
        for k in range(L.lpr_text_cut_count(h)):
            a, b = np.zeros(4096), C.c_double()
            na = L.lpr_text_cut(h, k, a.ctypes.data_as(C.POINTER(C.c_double)), C.byref(b))
            cuts.append(dict(a=a[:na].copy(), b=b.value))
        out["cuts"] = cuts
        return out
    finally:
        L.lpr_text_free(h)


def _masks(f_mask, f_len, h, chunks):
    """Per callback chunk: the bool[,] highlight (None for a null mask) and the chunk's length."""
    f_mask.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_long]
    f_len.argtypes = [C.c_void_p, C.c_int]
    f_len.restype = C.c_long
    masks, lens = [], []
    for k in range(chunks):
        r, c = C.c_int(), C.c_int()
        has = f_mask(h, k, C.byref(r), C.byref(c), None, 0)
        mk = None
        if has == 1:
            mk = np.zeros((r.value, c.value), dtype=np.uint8)
            f_mask(h, k, C.byref(r), C.byref(c), mk.ctypes.data_as(C.c_void_p), mk.size)
        masks.append(mk)
        lens.append(f_len(h, k))
    return masks, lens


def parse_text(text):
    L = lib()
    sense, m, n = C.c_int(), C.c_int(), C.c_int()
    err = C.create_string_buffer(512)
    rc = L.lpr_parse_text(text.encode("utf-8"), C.byref(sense), C.byref(m), C.byref(n), None, None, None, None, err, 512)
    if rc != 0:
        raise ValueError(err.value.decode("utf-8"))
    A = np.zeros((m.value, n.value))
    rel = np.zeros(m.value, dtype=np.int32)
    b = np.zeros(m.value)
    c = np.zeros(n.value)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.lpr_parse_text(text.encode("utf-8"), C.byref(sense), C.byref(m), C.byref(n), A.ctypes.data_as(dp),
                     rel.ctypes.data_as(ip), b.ctypes.data_as(dp), c.ctypes.data_as(dp), err, 512)
    return dict(sense=sense.value, A=A, rel=rel, b=b, c=c)


def fmt_custom(v, d=3):
    return lib().lpr_fmt_custom(float(v), d).decode("utf-8")


def fmt_fixed(v, d=3):
    return lib().lpr_fmt_fixed(float(v), d).decode("utf-8")


def fmt_roundtrip(v):
    return lib().lpr_fmt_roundtrip(float(v)).decode("utf-8")


def math_round(v, d):
    return lib().lpr_math_round(float(v), d)


def normalize_key(s):
    return lib().lpr_normalize_key(s.encode("utf-8")).decode("utf-8")

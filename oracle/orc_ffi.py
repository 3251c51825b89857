"""ctypes bindings for oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the product package.  See oracle/oracle.h for what each call restates.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_lp = C.POINTER(C.c_long)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_last_error.restype = C.c_char_p
        L.orc_primal_batch.restype = C.c_long
        L.orc_solve_text.restype = C.c_void_p
        for f in ("orc_text_error", "orc_text_log", "orc_text_report", "orc_text_summary"):
            getattr(L, f).restype = C.c_char_p
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_text_code.argtypes = [C.c_void_p]
        L.orc_text_masks.argtypes = [C.c_void_p]
        L.orc_text_free.argtypes = [C.c_void_p]
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_text_result_dims.argtypes = [C.c_void_p, ip, ip, ip, ip]
        L.orc_text_tableau.restype = dp
        L.orc_text_solution.restype = dp
        L.orc_text_basis.restype = ip
        L.orc_text_z.restype = C.c_double
        for f in ("orc_text_tableau", "orc_text_solution", "orc_text_basis", "orc_text_z", "orc_text_cut_count",
                  "orc_text_cut_end"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_text_cut.argtypes = [C.c_void_p, C.c_int, ip, ip, dp, dp]
        for f in ("orc_fmt_custom", "orc_fmt_fixed"):
            getattr(L, f).restype = C.c_char_p
            getattr(L, f).argtypes = [C.c_double, C.c_int]
        L.orc_fmt_roundtrip.restype = C.c_char_p
        L.orc_fmt_roundtrip.argtypes = [C.c_double]
        L.orc_math_round.restype = C.c_double
        L.orc_math_round.argtypes = [C.c_double, C.c_int]
        _LIB = L
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def last_error():
    return lib().orc_last_error().decode("utf-8")


def fmt_custom(v, decimals=3):
    return lib().orc_fmt_custom(float(v), decimals).decode("utf-8")


def fmt_fixed(v, decimals=3):
    return lib().orc_fmt_fixed(float(v), decimals).decode("utf-8")


def fmt_roundtrip(v):
    return lib().orc_fmt_roundtrip(float(v)).decode("utf-8")


def math_round(v, digits):
    return lib().orc_math_round(float(v), digits)


def parse_text(text):
    L = lib()
    sense, m, n = C.c_int(), C.c_int(), C.c_int()
    rc = L.orc_parse_text(text.encode("utf-8"), C.byref(sense), C.byref(m), C.byref(n), None, None, None, None)
    if rc != 0:
        raise ValueError((rc, last_error()))
    A = np.zeros((m.value, n.value))
    rel = np.zeros(m.value, dtype=np.int32)
    b = np.zeros(m.value)
    c = np.zeros(n.value)
    rc = L.orc_parse_text(text.encode("utf-8"), C.byref(sense), C.byref(m), C.byref(n), _d(A), _i(rel), _d(b), _d(c))
    if rc != 0:
        raise ValueError((rc, last_error()))
    return dict(sense=sense.value, A=A, rel=rel, b=b, c=c)


def tableau_dims(m, n, rel):
    rows, cols = C.c_int(), C.c_int()
    rel = np.ascontiguousarray(rel, dtype=np.int32)
    lib().orc_tableau_dims(m, n, _i(rel), C.byref(rows), C.byref(cols))
    return rows.value, cols.value


def _prep(A, rel, b, c):
    A = np.ascontiguousarray(A, dtype=np.float64)
    m, n = A.shape
    rel = np.zeros(m, dtype=np.int32) if rel is None else np.ascontiguousarray(rel, dtype=np.int32)
    b = np.ascontiguousarray(b, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    return A, rel, b, c, m, n


def primal_solve(A, b, c, rel=None, sense=0, max_iterations=10000, history=False, pivots_cap=None):
    A, rel, b, c, m, n = _prep(A, rel, b, c)
    rows, cols = tableau_dims(m, n, rel)
    cap = pivots_cap if pivots_cap is not None else max_iterations
    status, npiv = C.c_int(), C.c_int()
    pivots = np.full((cap, 2), -1, dtype=np.int32)
    basis = np.zeros(rows - 1, dtype=np.int32)
    x = np.zeros(n)
    z = C.c_double()
    T = np.zeros((rows, cols))
    hist_cap = 0
    hist = None
    if history:
        hist_cap = history if isinstance(history, int) and history is not True else 512
        hist = np.zeros((hist_cap, rows, cols))
    rc = lib().orc_primal_solve(m, n, sense, _d(A), _i(rel), _d(b), _d(c), max_iterations, C.byref(status),
                                C.byref(npiv), _i(pivots), cap, _i(basis), _d(x), C.byref(z), _d(T), _d(hist),
                                hist_cap)
    out = dict(rc=rc, status=status.value, n_pivots=npiv.value, pivots=pivots[: min(npiv.value, cap)], basis=basis,
               x=x, z=z.value, tableau=T, rows=rows, cols=cols)
    if rc != 0:
        out["error"] = last_error()
    if history:
        out["history"] = hist[: min(npiv.value + 1, hist_cap)]
    return out


def dual_solve(A, b, c, rel=None, sense=0, history=False, pivots_cap=10200):
    A, rel, b, c, m, n = _prep(A, rel, b, c)
    rows, cols = tableau_dims(m, n, rel)
    status, npiv, silent = C.c_int(), C.c_int(), C.c_int()
    pivots = np.full((pivots_cap, 2), -1, dtype=np.int32)
    basis = np.zeros(rows - 1, dtype=np.int32)
    x = np.zeros(n)
    z = C.c_double()
    T = np.zeros((rows, cols))
    hist_cap = 0
    hist = None
    if history:
        hist_cap = 512
        hist = np.zeros((hist_cap, rows, cols))
    rc = lib().orc_dual_solve(m, n, sense, _d(A), _i(rel), _d(b), _d(c), C.byref(status), C.byref(npiv),
                              C.byref(silent), _i(pivots), pivots_cap, _i(basis), _d(x), C.byref(z), _d(T),
                              _d(hist), hist_cap)
    out = dict(rc=rc, status=status.value, n_pivots=npiv.value, silent=silent.value,
               pivots=pivots[: min(npiv.value, pivots_cap)], basis=basis, x=x, z=z.value, tableau=T, rows=rows,
               cols=cols)
    if rc != 0:
        out["error"] = last_error()
    if history:
        out["history"] = hist[: min(npiv.value - silent.value + 1, hist_cap)]
    return out


def primal_batch(A, b, c, max_iterations=10000, threads=1, with_format=False, want_tableau=False):
    A = np.ascontiguousarray(A, dtype=np.float64)
    count, m, n = A.shape
    b = np.ascontiguousarray(b, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    status = np.zeros(count, dtype=np.int32)
    npiv = np.zeros(count, dtype=np.int32)
    basis = np.zeros((count, m), dtype=np.int32)
    x = np.zeros((count, n))
    z = np.zeros(count)
    T = np.zeros((count, m + 1, n + m + 1)) if want_tableau else None
    total = lib().orc_primal_batch(count, m, n, _d(A), _d(b), _d(c), max_iterations, threads, int(with_format),
                                   _i(status), _i(npiv), _i(basis), _d(x), _d(z), _d(T))
    return dict(total_pivots=total, status=status, n_pivots=npiv, basis=basis, x=x, z=z, tableau=T)


def primal_core(T, basis, max_pivots, pivots_cap=None):
    """In-place arithmetic pivots on a prebuilt tableau (rows = m+1)."""
    assert T.flags.c_contiguous and T.dtype == np.float64
    rows, width = T.shape
    cap = pivots_cap if pivots_cap is not None else max_pivots
    piv = np.full((cap, 2), -1, dtype=np.int32)
    npiv = C.c_int()
    st = lib().orc_primal_core(_d(T), rows - 1, width, _i(basis), max_pivots, C.byref(npiv), _i(piv), cap)
    return st, npiv.value, piv[: min(npiv.value, cap)]


def bnb_simplex(A, b, c, rel=None, sense=0, node_cap=4096):
    A, rel, b, c, m, n = _prep(A, rel, b, c)
    found, nn = C.c_int(), C.c_int()
    best_z = C.c_double()
    tp = C.c_long()
    best_x = np.zeros(n)
    oc = np.zeros(node_cap, dtype=np.int32)
    al = np.zeros(node_cap, dtype=np.int32)
    pv = np.zeros(node_cap, dtype=np.int32)
    nz = np.zeros(node_cap)
    bv = np.zeros(node_cap, dtype=np.int32)
    dp = np.zeros(node_cap, dtype=np.int32)
    rc = lib().orc_bnb_simplex(m, n, sense, _d(A), _i(rel), _d(b), _d(c), C.byref(found), C.byref(best_z), _d(best_x),
                               C.byref(nn), C.byref(tp), node_cap, _i(oc), _i(al), _i(pv), _d(nz), _i(bv), _i(dp))
    k = min(nn.value, node_cap)
    return dict(rc=rc, found=bool(found.value), best_z=best_z.value, best_x=best_x, n_nodes=nn.value,
                total_pivots=tp.value, outcome=oc[:k], algo=al[:k], pivots=pv[:k], z=nz[:k], branch_var=bv[:k],
                depth=dp[:k])


def bnb_pooled(A, b, c, rel=None, sense=0, batch=64, node_cap=1 << 20, max_nodes=0):
    """Mode B, the pooled tree (NOT the reference's tree; see oracle/orc_pooled.cpp)."""
    A, rel, b, c, m, n = _prep(A, rel, b, c)
    found = C.c_int()
    best_z = C.c_double()
    nn, tp, rounds, skipped = C.c_long(), C.c_long(), C.c_long(), C.c_long()
    best_x = np.zeros(n)
    nid = np.zeros(node_cap, dtype=np.int32)
    oc = np.zeros(node_cap, dtype=np.int32)
    pv = np.zeros(node_cap, dtype=np.int32)
    nz = np.zeros(node_cap)
    L = lib()
    L.orc_bnb_pooled.argtypes = [C.c_int, C.c_int, C.c_int, c_dp, c_ip, c_dp, c_dp, C.c_int, C.c_long, c_ip, c_dp, c_dp, c_lp, c_lp,
                                 c_lp, c_lp, C.c_long, c_ip, c_ip, c_ip, c_dp]
    rc = L.orc_bnb_pooled(m, n, sense, _d(A), _i(rel), _d(b), _d(c), batch, max_nodes, C.byref(found), C.byref(best_z), _d(best_x),
                          C.byref(nn), C.byref(tp), C.byref(rounds), C.byref(skipped), node_cap, _i(nid), _i(oc), _i(pv),
                          _d(nz))
    k = min(nn.value, node_cap)
    return dict(rc=rc, found=bool(found.value), best_z=best_z.value, best_x=best_x, n_nodes=nn.value,
                total_pivots=tp.value, rounds=rounds.value, skipped=skipped.value, node_id=nid[:k], outcome=oc[:k],
                pivots=pv[:k], z=nz[:k])


def knapsack(profit, weight, capacity, eval_cap=1 << 22):
    p = np.ascontiguousarray(profit, dtype=np.float64)
    w = np.ascontiguousarray(weight, dtype=np.float64)
    n = p.shape[0]
    found = C.c_int()
    best = C.c_double()
    ne, npops = C.c_long(), C.c_long()
    bx = np.zeros(n, dtype=np.int32)
    par = np.zeros(eval_cap, dtype=np.int32)
    ch = np.zeros(eval_cap, dtype=np.int32)
    var = np.zeros(eval_cap, dtype=np.int32)
    bd = np.zeros(eval_cap)
    wt = np.zeros(eval_cap)
    fr = np.zeros(eval_cap, dtype=np.int32)
    dc = np.zeros(eval_cap, dtype=np.int32)
    L = lib()
    L.orc_knapsack.argtypes = [C.c_int, c_dp, c_dp, C.c_double, c_ip, c_dp, c_ip, c_lp, c_lp, C.c_long, c_ip, c_ip,
                               c_ip, c_dp, c_dp, c_ip, c_ip]
    rc = L.orc_knapsack(n, _d(p), _d(w), float(capacity), C.byref(found), C.byref(best), _i(bx), C.byref(ne),
                        C.byref(npops), eval_cap, _i(par), _i(ch), _i(var), _d(bd), _d(wt), _i(fr), _i(dc))
    k = min(ne.value, eval_cap)
    return dict(rc=rc, found=bool(found.value), best=best.value, best_x=bx, n_evals=ne.value, n_pops=npops.value,
                parent=par[:k], child=ch[:k], var=var[:k], bound=bd[:k], weight=wt[:k], frac=fr[:k], decision=dc[:k])


def revised_solve(A, b, c, rel=None, sense=0, max_iterations=10000, cap=16384):
    A, rel, b, c, m, n = _prep(A, rel, b, c)
    L = lib()
    status, n_iters, z = C.c_int(), C.c_int(), C.c_double()
    enter, leave = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
    theta = np.zeros(cap)
    basis = np.zeros(m, dtype=np.int32)
    xB, Binv, x = np.zeros(m), np.zeros((m, m)), np.zeros(n)
    rc = L.orc_revised_solve(m, n, sense, _d(A), _i(rel), _d(b), _d(c), max_iterations, C.byref(status),
                             C.byref(n_iters), _i(enter), _i(leave), _d(theta), cap, _i(basis), _d(xB), _d(Binv), _d(x),
                             C.byref(z))
    if rc != 0:
        raise RuntimeError(last_error())
    k = min(n_iters.value, cap)
    return dict(status=status.value, n_iters=n_iters.value, enter=enter[:k], leave=leave[:k], theta=theta[:k],
                basis=basis, xB=xB, Binv=Binv, x=x, z=z.value)


def _masks(f_mask, f_len, h, chunks):
    """Per callback chunk: the bool[,] highlight (None for a null mask) and the chunk's length."""
    f_mask.argtypes = [C.c_void_p, C.c_int, c_ip, c_ip, C.c_void_p, C.c_long]
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


def solve_text(text, algorithm):
    L = lib()
    h = L.orc_solve_text(text.encode("utf-8"), algorithm.encode("utf-8"))
    try:
        out = dict(code=L.orc_text_code(h), error=L.orc_text_error(h).decode("utf-8"),
                   log=L.orc_text_log(h).decode("utf-8"), report=L.orc_text_report(h).decode("utf-8"),
                   summary=L.orc_text_summary(h).decode("utf-8"), chunks=L.orc_text_masks(h))
        rows, cols, nx, nb = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.orc_text_result_dims(h, C.byref(rows), C.byref(cols), C.byref(nx), C.byref(nb))
        out["tableau"] = (np.ctypeslib.as_array(L.orc_text_tableau(h), (rows.value, cols.value)).copy()
                          if rows.value else None)
        out["x"] = np.ctypeslib.as_array(L.orc_text_solution(h), (nx.value,)).copy() if nx.value else None
        out["basis"] = np.ctypeslib.as_array(L.orc_text_basis(h), (nb.value,)).copy() if nb.value else None
        out["z"] = L.orc_text_z(h)
        out["masks"], out["chunk_len"] = _masks(L.orc_text_mask, L.orc_text_chunk_len, h, out["chunks"])
        cuts = []
        n = max(nx.value, 1024)
        for k in range(L.orc_text_cut_count(h)):
            fv, row, b = C.c_int(), C.c_int(), C.c_double()
            a = np.zeros(n)
            na = L.orc_text_cut(h, k, C.byref(fv), C.byref(row), _d(a), C.byref(b))
            cuts.append(dict(frac_var=fv.value, row=row.value, a=a[:na].copy(), b=b.value))
        out["cuts"] = cuts
        out["cut_end"] = L.orc_text_cut_end(h)
        return out
    finally:
        L.orc_text_free(h)

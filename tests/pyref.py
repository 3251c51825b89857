"""Second, independent restatement of the reference's algorithms in plain Python loops (IEEE
doubles, no FMA), written straight from the C# sources.  Small cases only.  It exists to pin the
C++ oracle: tests/golden/*.json is produced by this file (tests/golden/make_golden.py) and the
oracle, the CUDA path and this file must all agree bit for bit.

  primal()    R/Models/PrimalSimplex.cs:57-127, 161-257
  dual()      R/Models/DualSimplex.cs:15-114, 117-158, 195-246
  bnb()       R/Models/Branch&Bound.cs:30-258
  knapsack()  R/Models/BranchAndBoundKnapsack.cs:58-328, 431-547
"""
import math

EPS = 1e-9
LE, GE, EQ = 0, 1, 2


class SolveError(Exception):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def _build(rows_a, rows_b, c):
    m, n = len(rows_a), len(c)
    w = n + m + 1
    T = [[0.0] * w for _ in range(m + 1)]
    for i in range(m):
        for j in range(n):
            T[i][j] = rows_a[i][j]
        T[i][n + i] = 1.0
        T[i][n + m] = rows_b[i]
    for j in range(n):
        T[m][j] = -c[j]
    return T, list(range(n, n + m))


def _entering(T):
    m = len(T) - 1
    best, mv = -1, -EPS
    for j in range(len(T[0]) - 1):
        if T[m][j] < mv:
            mv, best = T[m][j], j
    return best


def _leaving(T, e, margin):
    m = len(T) - 1
    best, row = math.inf, -1
    for i in range(m):
        a = T[i][e]
        if a > EPS:
            r = T[i][-1] / a
            if r < best - margin:
                best, row = r, i
    return row


def _pivot(T, r, c):
    piv = T[r][c]
    T[r] = [v / piv for v in T[r]]
    for i in range(len(T)):
        if i == r:
            continue
        f = T[i][c]
        T[i] = [T[i][j] - f * T[r][j] for j in range(len(T[i]))]


def primal(A, b, c, rel=None, sense=0, max_iterations=10000, history=False):
    m = len(A)
    rel = rel if rel is not None else [LE] * m
    c = [-v for v in c] if sense == 1 else list(c)
    for i in range(m):
        if rel[i] == GE:
            raise SolveError(-1, "ge")
        if b[i] < -1e-9:
            raise SolveError(-2, "negrhs")
    ra, rb = [], []
    for i in range(m):
        ra.append(list(A[i]))
        rb.append(b[i])
        if rel[i] == EQ:
            ra.append([v * -1 for v in A[i]])
            rb.append(b[i] * -1)
    T, basis = _build(ra, rb, c)
    pivots, hist = [], [[row[:] for row in T]] if history else None
    it = 1
    status = 0
    while True:
        if it > max_iterations:
            raise SolveError(-3, "iter")
        e = _entering(T)
        if e == -1:
            break
        l = _leaving(T, e, 1e-9)
        if l == -1:
            status = 1
            break
        _pivot(T, l, e)
        basis[l] = e
        pivots.append((e, l))
        if history:
            hist.append([row[:] for row in T])
        it += 1
    n = len(c)
    x = [0.0] * n
    for i in range(len(T) - 1):
        if basis[i] < n:
            x[basis[i]] = T[i][-1]
    return dict(status=status, pivots=pivots, basis=basis, x=x, z=T[-1][-1], tableau=T, history=hist)


def dual(A, b, c, rel=None, sense=0):
    m = len(A)
    rel = rel if rel is not None else [LE] * m
    c = [-v for v in c] if sense == 1 else list(c)
    ra, rb = [], []
    for i in range(m):
        if rel[i] == EQ:
            ra.append(list(A[i]))
            rb.append(b[i])
            ra.append([v * -1 for v in A[i]])
            rb.append(-b[i])
        else:
            a, bb = list(A[i]), b[i]
            if rel[i] == GE:
                a = [v * -1 for v in a]
                bb = bb * -1
            if bb < -EPS:
                a = [v * -1 for v in a]
                bb = bb * -1
            ra.append(a)
            rb.append(bb)
    T, basis = _build(ra, rb, c)
    pivots = []
    silent = 0
    for _ in range(100):
        e = _entering(T)
        if e == -1:
            break
        l = _leaving(T, e, 1e-12)
        if l == -1:
            break
        _pivot(T, l, e)
        basis[l] = e
        pivots.append((e, l))
        silent += 1
    mm = len(T) - 1
    ns = len(T[0]) - 1
    it = 1
    while True:
        if it > 10000:
            raise SolveError(-3, "iter")
        leave, mn = -1, -EPS
        for i in range(mm):
            if T[i][ns] < mn:
                mn, leave = T[i][ns], i
        if leave == -1:
            status = 0
            break
        enter, best = -1, math.inf
        for j in range(ns):
            a = T[leave][j]
            if a < -EPS:
                r = T[mm][j] / (-a)
                if r < best - 1e-12:
                    best, enter = r, j
        if enter == -1:
            status = 2
            break
        _pivot(T, leave, enter)
        basis[leave] = enter
        pivots.append((enter, leave))
        it += 1
    n = len(c)
    x = [0.0] * n
    for i in range(mm):
        if basis[i] < n:
            x[basis[i]] = T[i][-1]
    return dict(status=status, pivots=pivots, silent=silent, basis=basis, x=x, z=T[-1][-1], tableau=T)


def _round_even(v):
    return float(round(v))  # Python round() is ties-to-even like Math.Round


def bnb(A, b, c, rel=None, sense=0):
    """Returns (found, best_z, best_x, node outcomes in solve order)."""
    BB = 1e-6
    n = len(c)
    m0 = len(A)
    rel0 = list(rel) if rel is not None else [LE] * m0
    state = dict(best=-math.inf, best_x=None, counter=1, nodes=[])

    def algo(rl):
        return 1 if any(r in (GE, EQ) for r in rl) else 0

    def feasible(x, Ar, rl, br):
        for a, r, bb in zip(Ar, rl, br):
            s = 0.0
            for i in range(n):
                s += a[i] * x[i]
            if r == LE and s > bb + BB:
                return False
            if r == GE and s < bb - BB:
                return False
            if r == EQ and abs(s - bb) > BB:
                return False
        return all(v >= -BB for v in x)

    def integral(x):
        return all(abs(v - _round_even(v)) <= BB for v in x)

    def solve_lp(Ar, rl, br):
        if algo(rl) == 1:
            try:
                dual(Ar, br, c, rl, sense)
            except SolveError:
                return "error", None, None
            return "invalid", None, None
        try:
            r = primal(Ar, br, c, rl, sense)
        except SolveError:
            return "error", None, None
        return "ok", r["x"], r["z"]

    def node(Ar, rl, br, depth):
        if depth > 200:
            state["nodes"].append(7)
            return
        st, x, z = solve_lp(Ar, rl, br)
        if st == "error":
            state["nodes"].append(0)
            return
        if st == "invalid":
            state["nodes"].append(1)
            return
        if not feasible(x, Ar, rl, br):
            state["nodes"].append(2)
            return
        if z <= state["best"] + BB:
            state["nodes"].append(3)
            return
        if integral(x):
            state["best"] = z
            state["best_x"] = [_round_even(v) for v in x]
            state["nodes"].append(4)
            return
        fi, md = -1, float("inf")
        for i in range(n):
            fp = x[i] - math.floor(x[i])
            if fp > BB and (1 - fp) > BB:
                d = abs(fp - 0.5)
                if d < md or (d == md and i < fi):
                    md, fi = d, i
        if fi == -1:
            state["nodes"].append(6)
            return
        fl, ce = int(math.floor(x[fi])), int(math.ceil(x[fi]))
        state["nodes"].append(5)
        unit = [0.0] * n
        unit[fi] = 1.0
        node(Ar + [unit], rl + [GE], br + [float(ce)], depth + 1)
        node(Ar + [unit], rl + [LE], br + [float(fl)], depth + 1)

    Ar = [list(r) for r in A]
    st, x, z = solve_lp(Ar, rel0, list(b))
    if st != "ok":
        state["nodes"].append(0 if st == "error" else 1)
        return False, -math.inf, None, state["nodes"]
    if integral(x) and feasible(x, Ar, rel0, list(b)):
        state["nodes"].append(4)
        return True, z, [_round_even(v) for v in x], state["nodes"]
    state["nodes"].append(5)
    node(Ar, rel0, list(b), 0)
    return state["best_x"] is not None, state["best"], state["best_x"], state["nodes"]


def knapsack(p, w, cap):
    """Returns (found, best, best_x, evals) with evals = [(bound, weight, frac_rank, decision), ...]."""
    n = len(p)
    KE = 1e-9

    def ratio(i):
        return p[i] / w[i] if w[i] > 0 else math.inf

    order = sorted(range(n), key=lambda i: (-ratio(i), -p[i]))  # stable
    # python's sort with a tuple key is stable and matches OrderByDescending.ThenByDescending
    rank_of = {o: s for s, o in enumerate(order)}

    def relax(assigned):
        relaxed = [0.0] * n
        weight = profit = 0.0
        frac = -1
        for i in range(n):
            if assigned[i] == 1:
                relaxed[i] = 1.0
                weight += w[i]
                profit += p[i]
        if weight > cap + KE:
            return relaxed, profit, -1, weight
        for s in range(n):
            o = order[s]
            if assigned[o] == 1:
                continue
            if assigned[o] == 0:
                continue
            if weight + w[o] <= cap + KE:
                relaxed[o] = 1.0
                weight += w[o]
                profit += p[o]
            else:
                remain = cap - weight
                if remain > KE and w[o] > KE:
                    f = remain / w[o]
                    relaxed[o] = f
                    profit += p[o] * f
                    weight += w[o] * f
                    frac = s
                break
        return relaxed, profit, frac, weight

    heap = []

    def push(item):
        heap.append(item)
        ci = len(heap) - 1
        while ci > 0:
            pi = (ci - 1) // 2
            if not heap[ci][0] > heap[pi][0]:
                break
            heap[ci], heap[pi] = heap[pi], heap[ci]
            ci = pi

    def pop():
        li = len(heap) - 1
        heap[0], heap[li] = heap[li], heap[0]
        ret = heap.pop()
        li = len(heap) - 1
        i = 0
        while True:
            l, r, big = 2 * i + 1, 2 * i + 2, i
            if l <= li and heap[l][0] > heap[big][0]:
                big = l
            if r <= li and heap[r][0] > heap[big][0]:
                big = r
            if big == i:
                break
            heap[i], heap[big] = heap[big], heap[i]
            i = big
        return ret

    best, best_x = -math.inf, [0] * n
    evals = []
    root = [-1] * n
    rr = relax(root)
    evals.append((rr[1], rr[3], rr[2], 0))
    push((rr[1], root))
    pops = 0
    while heap:
        bound, assigned = pop()
        pops += 1
        if bound <= best + KE:
            continue
        relaxed, prof, frac, wt = relax(assigned)
        if frac == -1:
            if wt <= cap + KE and prof > best + KE:
                best = prof
                best_x = [1 if v >= 0.5 else 0 for v in relaxed]
            continue
        o = order[frac]
        for side in (0, 1):
            ch = assigned[:]
            ch[o] = side
            rl, pr, fr, cw = relax(ch)
            if cw > cap + KE:
                dec = 1
            elif pr > best + KE:
                if all(abs(v - _round_even(v)) < KE for v in rl):
                    dec = 2
                    if pr > best + KE:
                        best = pr
                        best_x = [int(_round_even(v)) for v in rl]
                else:
                    dec = 3
                    push((pr, ch))
            else:
                dec = 4
            evals.append((pr, cw, fr, dec))
    return best > -math.inf, best, best_x, evals, pops, order


def cutting_plane(A, b, c, rel=None, sense=0):
    """CuttingPlane.Solve (R/Models/CuttingPlane.cs:13-164): <= 50 rounds of primal + one cut read
    from tableau row `basis position + 1` (the reference's z-row-first assumption, :113)."""
    import math
    n = len(c)
    A = [list(r) for r in A]
    b = list(b)
    rel = list(rel) if rel is not None else [LE] * len(A)
    cuts, rounds = [], []
    for _ in range(50):
        try:
            r = primal(A, b, c, rel, sense)
        except SolveError as e:
            return dict(end=2, code=e.code, cuts=cuts, rounds=rounds)
        rounds.append((len(r["pivots"]), r["status"]))
        x = r["x"][:n]
        frac = -1
        for i, v in enumerate(x):
            f = v - math.floor(v)
            if 1e-9 < f < 1 - 1e-9:
                frac = i
                break
        if frac == -1:
            return dict(end=0, cuts=cuts, rounds=rounds, x=x, z=r["z"], tableau=r["tableau"], basis=r["basis"])
        row = -1
        for i, bv in enumerate(r["basis"]):
            if bv == frac:
                row = i + 1
                break
        if row == -1:
            return dict(end=3, cuts=cuts, rounds=rounds)
        trow = r["tableau"][row]
        a = [0.0] * n
        for j in range(n):
            fj = trow[j] - math.floor(trow[j])
            if fj > 1e-9:
                a[j] = fj
        f0 = trow[-1] - math.floor(trow[-1])
        cuts.append((frac, row, a, f0))
        A.append(a)
        b.append(f0)
        rel.append(LE)
    return dict(end=1, cuts=cuts, rounds=rounds)


def _invert(M):
    """RevisedPrimalSimplex.Invert (R/Models/RevisedPrimalSimplex.cs:409-456)."""
    n = len(M)
    A = [list(M[i]) + [1.0 if j == i else 0.0 for j in range(n)] for i in range(n)]
    for col in range(n):
        piv_row, best = col, abs(A[col][col])
        for r in range(col + 1, n):
            if abs(A[r][col]) > best:
                best, piv_row = abs(A[r][col]), r
        if abs(A[piv_row][col]) < 1e-9:
            raise SolveError(-11, "singular")
        if piv_row != col:
            A[col], A[piv_row] = A[piv_row], A[col]
        piv = A[col][col]
        A[col] = [v / piv for v in A[col]]
        for r in range(n):
            if r == col:
                continue
            f = A[r][col]
            A[r] = [A[r][j] - f * A[col][j] for j in range(2 * n)]
    return [row[n:] for row in A]


def _seqdot(a, b):
    s = 0.0
    for u, v in zip(a, b):
        s += u * v
    return s


def revised(A, b, c, rel=None, sense=0, max_iterations=10000):
    """RevisedPrimalSimplex.Solve (R/Models/RevisedPrimalSimplex.cs:17-145): numerics only."""
    m, n = len(A), len(c)
    rel = rel if rel is not None else [LE] * m
    if not all(rel[i] == LE and b[i] >= -1e-9 for i in range(m)):
        raise SolveError(-10, "unsupported")
    cc = [-v for v in c] if sense == 0 else list(c)   # Standardize negates C for MAX
    ntot = n + m
    Af = [list(A[i]) + [1.0 if j == i else 0.0 for j in range(m)] for i in range(m)]
    cf = cc + [0.0] * m
    Bidx, Nidx = list(range(n, ntot)), list(range(n))

    def state():
        Binv = _invert([[Af[i][k] for k in Bidx] for i in range(m)])
        xB = [_seqdot(Binv[i], b) for i in range(m)]
        cB = [cf[k] for k in Bidx]
        return Binv, xB, cB
    Binv, xB, cB = state()
    pivots = []
    status = None
    for _ in range(max_iterations):
        piT = [_seqdot(cB, [Binv[i][j] for i in range(m)]) for j in range(m)]
        rN = [cf[k] - _seqdot(piT, [Af[i][k] for i in range(m)]) for k in Nidx]
        pos, best = -1, -1e-9
        for j, v in enumerate(rN):
            if v < best:
                best, pos = v, j
        if pos == -1:
            status = 0
            break
        e = Nidx[pos]
        d = [_seqdot(Binv[i], [Af[j][e] for j in range(m)]) for i in range(m)]
        row, theta_best = -1, float("inf")
        for i in range(m):
            if d[i] > 1e-9:
                theta = xB[i] / d[i]
                if theta < theta_best - 1e-12:
                    theta_best, row = theta, i
        if row == -1:
            status = 1
            break
        leaving = Bidx[row]
        Bidx[row] = e
        Nidx.pop(pos)
        Nidx.append(leaving)
        Binv, xB, cB = state()
        pivots.append((e, row, theta_best))
    if status is None:
        status = -3
    x = [0.0] * n
    for i, k in enumerate(Bidx):
        if k < n:
            x[k] = xB[i]
    return dict(status=status, pivots=pivots, basis=Bidx, xB=xB, Binv=Binv, x=x, z=_seqdot(c, x))

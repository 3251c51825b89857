"""Kernel-variant timing probe (development aid, run under gpurun).  CUDA events on torch's
current stream for the batched kernels, on the session stream for the streaming kernels."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linear_programming_solver_lpr381_b200 import _ffi as F  # noqa: E402
from linear_programming_solver_lpr381_b200 import api, workloads  # noqa: E402

F.check(F.lib().lpx_init(0))
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream()


def time_batched(kernel, threads=0, reg_variant=0, count=4096, m=64, n=128, reps=5, max_iterations=10000):
    A, b, c = workloads.batch_c2(count=count, m=m, n=n, seed=1)
    dA, db, dc = (torch.from_numpy(v).to(dev) for v in (A, b, c))
    st = torch.zeros(count, dtype=torch.int32, device=dev)
    npv = torch.zeros(count, dtype=torch.int32, device=dev)
    basis = torch.zeros((count, m), dtype=torch.int32, device=dev)
    x = torch.zeros((count, n), dtype=torch.float64, device=dev)
    z = torch.zeros(count, dtype=torch.float64, device=dev)
    T = torch.zeros((count, m + 1, n + m + 1), dtype=torch.float64, device=dev)
    tot = torch.zeros(1, dtype=torch.int64, device=dev)

    def launch():
        api.primal_solve_batched_dev(count, m, n, 0, dA.data_ptr(), None, db.data_ptr(), dc.data_ptr(), st.data_ptr(),
                                     npv.data_ptr(), basis.data_ptr(), x.data_ptr(), z.data_ptr(), T.data_ptr(),
                                     tot.data_ptr(), stream.cuda_stream, kernel=kernel, threads=threads,
                                     reg_variant=reg_variant, max_iterations=max_iterations)
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    tot.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        launch()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    piv = int(tot.item()) / reps
    return dict(kernel=kernel, threads=threads, reg_variant=reg_variant, ms=ms, pivots=piv, mpivots_s=piv / ms / 1e3)


def time_large(single_cta_select, per=40, reps=3):
    A, b, c = workloads.large_c3()
    s = api.Session(A, b, c, max_iterations=1 << 30, single_cta_select=single_cta_select)
    ss = torch.cuda.ExternalStream(s.stream)
    s.step_async(per)
    s.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ss)
    s.step_async(per * reps)
    e1.record(ss)
    s.sync()
    us = e0.elapsed_time(e1) * 1e3 / (per * reps)
    bytes_pp = 2 * 8 * s.rows * s.cols
    s.close()
    return dict(single_cta_select=single_cta_select, us_per_pivot=us, gbs=bytes_pp / us / 1e3, frac=bytes_pp / us / 1e3 / 6555.2)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "batched"):
        for kw in (dict(kernel=F.KERNEL_CTA_SMEM, threads=256), dict(kernel=F.KERNEL_CTA_SMEM, threads=512),
                   dict(kernel=F.KERNEL_CTA_REG, reg_variant=1), dict(kernel=F.KERNEL_CTA_REG, reg_variant=2)):
            print(json.dumps(time_batched(**kw)), flush=True)
    if what in ("all", "large"):
        for sc in (1, 0):
            print(json.dumps(time_large(sc)), flush=True)
    if what == "prof_reg":
        print(json.dumps(time_batched(kernel=F.KERNEL_CTA_REG, reg_variant=1, count=592, reps=1)), flush=True)
    if what == "prof_smem":
        print(json.dumps(time_batched(kernel=F.KERNEL_CTA_SMEM, threads=256, count=592, reps=1)), flush=True)
    if what == "largeprof":
        import ctypes as C
        A, b, c = workloads.large_c3()
        for sc in (1, 0):
            s = api.Session(A, b, c, max_iterations=1 << 30, single_cta_select=sc)
            s.step(20)
            us = (C.c_double * 3)()
            F.lib().lpx_session_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
            F.check(F.lib().lpx_session_profile(s._h, 60, us))
            print(json.dumps(dict(single_cta_select=sc, prep_us=us[0], update_us=us[1], per_pivot_us=us[2])), flush=True)
            s.close()
    if what == "block":
        A, b, c = workloads.large_c3()
        for kb, proto in ((8, 0), (4, 0), (6, 0), (8, 4), (16, 4)):
            s = api.Session(A, b, c, max_iterations=1 << 30, kblock=kb, single_cta_select=proto)
            ss = torch.cuda.ExternalStream(s.stream)
            s.step(32)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            npv = 256
            e0.record(ss)
            s.step_async(npv)
            e1.record(ss)
            st, tot = s.sync()
            us = e0.elapsed_time(e1) * 1e3 / npv
            print(json.dumps(dict(kblock=kb, protocol=proto, us_per_pivot=us, pivots_s=1e6 / us, status=st, total=tot,
                                  x_roofline=805568528 / us / 1e3 / 6555.2)), flush=True)
            s.close()
    if what == "stamps":
        import ctypes as C
        A, b, c = workloads.large_c3()
        s = api.Session(A, b, c, max_iterations=1 << 30, kblock=8)
        s.step(24)
        for rep in range(3):
            s.step(8)
            out = (C.c_ulonglong * 8)()
            F.lib().lpx_session_debug_stamps.argtypes = [C.c_void_p, C.c_void_p]
            F.check(F.lib().lpx_session_debug_stamps(s._h, out))
            t = list(out)
            print(t, flush=True)
            print(json.dumps(dict(argmin_sync1=t[1] - t[0], column_sync2=t[2] - t[1], ratio_scan=t[3] - t[2],
                                  row=t[4] - t[3], rhs_sync3=t[5] - t[4], step=t[5] - t[0])), flush=True)
        s.close()
    if what == "block8":
        A, b, c = workloads.large_c3()
        s = api.Session(A, b, c, max_iterations=1 << 30, kblock=8)
        s.step(32)
        st, tot = s.step(16)
        print(st, tot)
        s.close()
    if what == "blockprof":
        import ctypes as C
        A, b, c = workloads.large_c3()
        for kb, pv in ((8, 0), (12, 0), (16, 0), (16, 1)):
            s = api.Session(A, b, c, max_iterations=1 << 30, kblock=kb, pass_variant=pv)
            s.step(32)
            us = (C.c_double * 3)()
            F.lib().lpx_session_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
            F.check(F.lib().lpx_session_profile(s._h, 12, us))
            print(json.dumps(dict(kblock=kb, pass_variant=pv, lookahead_us=us[0], pass_us=us[1], per_block_us=us[2],
                                  us_per_pivot=us[2] / kb)), flush=True)
            s.close()
    if what == "knap":
        import time
        ps, ws, caps = zip(*[workloads.knapsack_c5(seed=13 + k) for k in range(16)])
        p, w, cap = np.stack(ps), np.stack(ws), np.array(caps)
        api.bnb_knapsack_batched(p[:1], w[:1], cap[:1])
        for sn, sd in ((8, 3), (8, 5), (16, 4), (4, 6), (2, 8), (1, 10), (32, 2)):
            l0 = F.lib().lpx_kernel_launches()
            t0 = time.perf_counter()
            try:
                r = api.bnb_knapsack_batched(p, w, cap, spec_nodes=sn, spec_depth=sd)
            except Exception as ex:
                print(json.dumps(dict(spec_nodes=sn, spec_depth=sd, error=str(ex))), flush=True)
                continue
            dt = time.perf_counter() - t0
            print(json.dumps(dict(spec_nodes=sn, spec_depth=sd, nodes=int(r["n_evals"].sum()), seconds=dt,
                                  nodes_s=int(r["n_evals"].sum()) / dt, launches=F.lib().lpx_kernel_launches() - l0)),
                  flush=True)
    if what == "blockwin":
        A, b, c = workloads.large_c3()
        for warm, calls, per in ((32, 1, 256), (384, 1, 640), (384, 5, 128), (32, 8, 32), (1024, 1, 256)):
            s = api.Session(A, b, c, max_iterations=1 << 30)
            ss = torch.cuda.ExternalStream(s.stream)
            s.step(warm)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ss)
            for _ in range(calls):
                s.step_async(per)
            e1.record(ss)
            st, tot = s.sync()
            us = e0.elapsed_time(e1) * 1e3 / (calls * per)
            print(json.dumps(dict(warm=warm, calls=calls, per=per, us_per_pivot=us, pivots_s=1e6 / us, status=st)), flush=True)
            s.close()
    if what == "benchlike":
        A, b, c = workloads.large_c3()
        for mode in ("host", "devptr", "devptr_events"):
            if mode == "host":
                s = api.Session(A, b, c, max_iterations=1 << 30)
            else:
                dA, db, dc = (torch.from_numpy(v).to(dev) for v in (A, b, c))
                torch.cuda.synchronize()
                s = api.Session(dA.data_ptr(), db.data_ptr(), dc.data_ptr(), device_ptrs=True, m=4096, n=8192,
                                max_iterations=1 << 30)
                del dA
            ss = torch.cuda.ExternalStream(s.stream)
            for _ in range(3):
                s.step_async(128)
            s.sync()
            torch.cuda.synchronize()
            if mode == "devptr_events":
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
                for k in range(5):
                    ev[k][0].record(ss)
                    s.step_async(128)
                    ev[k][1].record(ss)
                st, tot = s.sync()
                ms = sum(a.elapsed_time(bb) for a, bb in ev)
            else:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(ss)
                for k in range(5):
                    s.step_async(128)
                e1.record(ss)
                st, tot = s.sync()
                ms = e0.elapsed_time(e1)
            print(json.dumps(dict(mode=mode, us_per_pivot=ms * 1e3 / 640, status=st)), flush=True)
            s.close()
    if what == "bnbrep":
        # repeat-call timing of the two B&B entry points: is the first full-size call paying for
        # workspace growth?
        import time
        cnt = 256
        As, bs, cs = zip(*[workloads.ip_c4(seed=11 + k) for k in range(cnt)])
        A, b, c = np.stack(As), np.stack(bs), np.stack(cs)
        api.bnb_simplex_batched(A[:1], b[:1], c[:1])
        for rep in range(4):
            t0 = time.perf_counter()
            r = api.bnb_simplex_batched(A, b, c)
            dt = time.perf_counter() - t0
            print(json.dumps(dict(call="bnb_simplex_batched", rep=rep, s=dt, nodes=int(r["n_nodes"].sum()),
                                  nodes_per_s=int(r["n_nodes"].sum()) / dt)), flush=True)
        ps, ws, caps = zip(*[workloads.knapsack_c5(seed=13 + k) for k in range(16)])
        p, w, cap = np.stack(ps), np.stack(ws), np.array(caps)
        api.bnb_knapsack_batched(p[:1], w[:1], cap[:1])
        for rep in range(4):
            t0 = time.perf_counter()
            r = api.bnb_knapsack_batched(p, w, cap)
            dt = time.perf_counter() - t0
            print(json.dumps(dict(call="bnb_knapsack_batched", rep=rep, s=dt, nodes=int(r["n_evals"].sum()),
                                  nodes_per_s=int(r["n_evals"].sum()) / dt)), flush=True)
    if what == "lat":
        import ctypes as C
        out = (C.c_double * 8)()
        F.check(F.lib().lpx_measure_latencies(out))
        names = ["DADD", "DMUL", "DDIV", "REDUX", "SHFL", "LDS_roundtrip", "BAR13", "DSETP_SEL"]
        print(json.dumps({k: round(v, 1) for k, v in zip(names, out)}), flush=True)
    if what == "regcounts":
        for rv in (1, 2):
            for cnt in (148, 296, 592, 1184, 4096):
                print(json.dumps(time_batched(kernel=F.KERNEL_CTA_REG, reg_variant=rv, count=cnt, reps=5)), flush=True)
    if what == "prof_reg2":
        print(json.dumps(time_batched(kernel=F.KERNEL_CTA_REG, reg_variant=2, count=1184, reps=1)), flush=True)
    if what == "kblock":
        # pipelined C3 session at several look-ahead block sizes
        A, b, c = workloads.large_c3()
        for kb in (8, 10, 12, 16):
            s = api.Session(A, b, c, max_iterations=1 << 30, kblock=kb)
            ss = torch.cuda.ExternalStream(s.stream)
            per = 240  # multiple of 8, 10, 12, 16
            for _ in range(2):
                s.step_async(per)
            s.sync()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ss)
            for k in range(3):
                s.step_async(per)
            e1.record(ss)
            st, tot = s.sync()
            us = e0.elapsed_time(e1) * 1e3 / (3 * per)
            s.close()
            print(json.dumps(dict(kblock=kb, us_per_pivot=us, pivots_per_s=1e6 / us, status=int(st))), flush=True)
    if what == "knapsweep":
        import time
        ps, ws, caps = zip(*[workloads.knapsack_c5(seed=13 + k) for k in range(16)])
        p, w, cap = np.stack(ps), np.stack(ws), np.array(caps)
        api.bnb_knapsack_batched(p, w, cap)
        for sn, sd in ((16, 4), (16, 2), (16, 1), (16, 3), (32, 2), (32, 1), (24, 2), (48, 2), (32, 3), (64, 2), (8, 2), (64, 1)):
            best = 1e9
            for rep in range(2):
                t0 = time.perf_counter()
                r = api.bnb_knapsack_batched(p, w, cap, spec_nodes=sn, spec_depth=sd)
                best = min(best, time.perf_counter() - t0)
            print(json.dumps(dict(spec_nodes=sn, spec_depth=sd, s=best, nodes_per_s=int(r["n_evals"].sum()) / best)), flush=True)
    if what == "smemshape":
        for (m, n, th) in ((64, 128, 256), (90, 120, 256), (90, 120, 512), (110, 120, 512), (110, 120, 256)):
            for cnt in (148, 1184):
                r = time_batched(kernel=F.KERNEL_CTA_SMEM, threads=th, count=cnt, m=m, n=n, reps=3)
                r["m"], r["n"], r["count"] = m, n, cnt
                r["us_per_pivot_per_cta_if_1_per_sm"] = r["ms"] * 1e3 / (r["pivots"] / cnt) if cnt == 148 else None
                print(json.dumps(r), flush=True)
    if what == "smemprof":
        print(json.dumps(time_batched(kernel=F.KERNEL_CTA_SMEM, threads=512, count=592, m=90, n=120, reps=1)), flush=True)
    if what == "knapweak":
        import time
        for kind in ("uncorrelated", "weak"):
            ps, ws, caps = zip(*[workloads.knapsack_c5(seed=13 + k, kind=kind) for k in range(16)])
            p, w, cap = np.stack(ps), np.stack(ws), np.array(caps)
            api.bnb_knapsack_batched(p, w, cap)
            best = 1e9
            for rep in range(3):
                t0 = time.perf_counter()
                r = api.bnb_knapsack_batched(p, w, cap)
                best = min(best, time.perf_counter() - t0)
            print(json.dumps(dict(kind=kind, s=best, nodes=int(r["n_evals"].sum()), nodes_per_s=int(r["n_evals"].sum()) / best,
                                  per_instance=r["n_evals"].tolist())), flush=True)
    if what == "knapsingle":
        import time
        for kind, seed in (("uncorrelated", 13), ("weak", 22)):
            p, w, cap = workloads.knapsack_c5(seed=seed, kind=kind)
            api.bnb_knapsack(p, w, cap)
            for sn, sd in ((16, 2), (32, 2), (64, 2), (128, 2), (64, 3), (128, 3), (256, 2)):
                best = 1e9
                for rep in range(2):
                    t0 = time.perf_counter()
                    r = api.bnb_knapsack(p, w, cap, spec_nodes=sn, spec_depth=sd)
                    best = min(best, time.perf_counter() - t0)
                print(json.dumps(dict(kind=kind, spec_nodes=sn, spec_depth=sd, s=best, nodes_per_s=r["n_evals"] / best)), flush=True)

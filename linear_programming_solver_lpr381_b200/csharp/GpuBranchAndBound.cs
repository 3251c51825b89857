// Drop-in ILPAlgorithm for "Branch and Bound" / "Branch and Bound Knapsack": the node relaxations
// run on the GPU (lpx_bnb_simplex / lpx_bnb_knapsack); the per-node records arrive through a host
// callback in the reference's order, so the log lines of Models/Branch&Bound.cs:128-258 can be
// produced unchanged from them.  The complete text replay lives in the C++ host layer
// (../host/branch_and_bound.cpp, ../host/knapsack.cpp); this file shows the managed binding.
using System;
using System.Runtime.InteropServices;

namespace Linear_Programming_Solver.Models
{
    public class GpuBranchAndBound : ILPAlgorithm
    {
        public SimplexResult Solve(LPProblem problem, Action<string, bool[,]> updatePivot = null)
        {
            int m = problem.Constraints.Count, n = problem.NumVars;
            var A = new double[m * n];
            var rel = new int[m];
            var b = new double[m];
            for (int i = 0; i < m; i++)
            {
                Array.Copy(problem.Constraints[i].A, 0, A, i * n, n);
                rel[i] = (int)problem.Constraints[i].Relation;
                b[i] = problem.Constraints[i].B;
            }
            var opt = new LpxOptions();
            LpxNative.lpx_default_options(ref opt);
            var bestX = new double[n];
            var replay = new BranchAndBoundLogReplay(problem, updatePivot);   // formats Log(...) lines per record
            LpxBnbNodeFn cb = (ref LpxBnbNode node, IntPtr user) => replay.OnNode(ref node);
            int rc = LpxNative.lpx_bnb_simplex(m, n, (int)problem.ObjectiveSense, A, rel, b, problem.C, ref opt,
                updatePivot != null ? 1 : 0, out int found, out double bestZ, bestX, out int nNodes,
                out long lpPivots, out int rootStatus, updatePivot != null ? cb : null, IntPtr.Zero);
            GC.KeepAlive(cb);
            if (rc != 0) throw new Exception(LpxNative.LastError());
            return replay.BuildReport(found != 0, bestZ, bestX, rootStatus);
        }
    }
}

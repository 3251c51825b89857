// Drop-in ILPAlgorithm for "Branch and Bound": same contract as BranchAndBound.Solve
// (Models/Branch&Bound.cs:30-123).  Every LP relaxation runs on the GPU (lpx_bnb_simplex); the node
// records arrive through a host callback in the reference's own order (root LP, then each SolveNode
// call, depth first, ceil child first), and the log lines of Models/Branch&Bound.cs:128-258 are
// re-emitted from them.  This is the C# twin of ../host/bnb_host.cpp (which is compiled and tested
// here); it is shipped as source because this image has no .NET toolchain.
using System;
using System.Collections.Generic;
using System.Globalization;
using System.Linq;
using System.Runtime.InteropServices;
using System.Text;

namespace Linear_Programming_Solver.Models
{
    public class GpuBranchAndBound : ILPAlgorithm
    {
        private const double EPS = 1e-6;
        private static string F3(double v) => v.ToString("F3", CultureInfo.InvariantCulture);
        private static string F6(double v) => v.ToString("F6", CultureInfo.InvariantCulture);

        private sealed class Rec { public int Parent, Var, Rel; public double Rhs; }

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
            void Log(string msg) => updatePivot?.Invoke(msg + Environment.NewLine, null);
            string RowText(double[] a, int off, int r, double rhs) =>
                string.Join(" + ", Enumerable.Range(0, n).Where(j => a[off + j] != 0)
                    .Select(j => $"{F3(a[off + j])}x{j + 1}")) + $" {(Rel)r} {F3(rhs)}";

            Log("=== Branch & Bound Algorithm ===");
            Log("Objective: Maximize " + string.Join(" + ", problem.C.Select((c, i) => $"{F3(c)}x{i + 1}")));
            Log("Subject to:");
            for (int i = 0; i < m; i++) Log(RowText(A, i * n, rel[i], b[i]));
            Log("x_j >= 0, integer");
            bool dualRoot = rel.Any(r => r != 0);
            Log($"Branch & Bound: Using {(dualRoot ? "Dual Simplex" : "Primal Simplex")} for the ROOT LP relaxation.");

            var recs = new List<Rec>();
            double best = double.NegativeInfinity;
            int counter = 1;
            double[,] rootTableau = null;
            int[] rootBasis = null;

            string Constraints(int rec)
            {
                var chain = new List<Rec>();
                for (int k = rec; k >= 0 && recs[k].Var >= 0; k = recs[k].Parent) chain.Add(recs[k]);
                chain.Reverse();
                var parts = Enumerable.Range(0, m).Select(i => RowText(A, i * n, rel[i], b[i])).ToList();
                foreach (var c in chain) parts.Add($"{F3(1.0)}x{c.Var + 1} {(Rel)c.Rel} {F3(c.Rhs)}");
                return string.Join("; ", parts);
            }

            void LpText(ref LpxBnbNode nd, int[] pivots, double[] x)
            {
                if (updatePivot == null || nd.n_history <= 0) return;
                int rows = nd.rows, cols = nd.cols;
                var names = Enumerable.Range(0, n).Select(j => $"x{j + 1}")
                    .Concat(Enumerable.Range(0, rows - 1).Select(j => $"c{j + 1}")).ToArray();
                var basis = Enumerable.Range(n, rows - 1).ToArray();
                for (int k = 0; k < nd.silent_pivots; k++) basis[pivots[2 * k + 1]] = pivots[2 * k];
                var hist = new double[nd.n_history * rows * cols];
                Marshal.Copy(nd.history, hist, 0, hist.Length);
                for (int k = 0; k < nd.n_history; k++)
                {
                    var T = new double[rows, cols];
                    Buffer.BlockCopy(hist, k * rows * cols * 8, T, 0, rows * cols * 8);
                    bool[,] hl = null;
                    if (k > 0)
                    {
                        int e = pivots[2 * (nd.silent_pivots + k - 1)], l = pivots[2 * (nd.silent_pivots + k - 1) + 1];
                        basis[l] = e;
                        hl = new bool[rows, cols];
                        for (int j = 0; j < cols; j++) hl[l, j] = true;
                        for (int i = 0; i < rows; i++) hl[i, e] = true;
                    }
                    var sb = new StringBuilder();
                    if (nd.algo == 1) DualSimplex.AppendTableau(sb, T, basis, names, k);   // made internal, see INTEGRATION.md
                    else PrimalSimplex.AppendTableau(sb, T, basis, names, k);
                    updatePivot(sb.ToString(), hl);
                    if (nd.index == 0 && k == nd.n_history - 1) { rootTableau = T; rootBasis = (int[])basis.Clone(); }
                }
            }

            LpxBnbNodeFn cb = (ref LpxBnbNode nd, IntPtr user) =>
            {
                var pivots = new int[2 * nd.n_pivots];
                if (nd.n_pivots > 0) Marshal.Copy(nd.pivots, pivots, 0, pivots.Length);
                var x = new double[n];
                if (nd.x != IntPtr.Zero) Marshal.Copy(nd.x, x, 0, n);
                var rec = new Rec { Parent = nd.parent, Var = nd.bound_var, Rel = nd.is_ceil_child, Rhs = nd.bound_val };
                if (nd.index == 0) rec.Var = -1;
                recs.Add(rec);
                string xs = string.Join(", ", x.Select(F3));
                if (nd.index == 0)
                {
                    LpText(ref nd, pivots, x);
                    if (nd.outcome == 0) { Log($"Root Problem: LP relaxation infeasible or error: {LpxNative.StatusMessage(nd.lp_status)}"); return; }
                    if (nd.outcome == 1) { Log("Root Problem: Invalid Simplex result (missing Solution, Tableau, Basis, or VarNames)."); return; }
                    Log($"Root Problem LP solution: z* = {F3(nd.z)}, x* = [{xs}]");
                    Log("Root Problem optimal tableau displayed above.");
                    if (nd.outcome == 4) { best = nd.z; Log("Root Problem is already integral and feasible. Branch & Bound not required."); }
                    else Log("Root solution is fractional → starting Branch & Bound.");
                    return;
                }
                var ids = new int[nd.id_path_len];
                if (nd.id_path_len > 0) Marshal.Copy(nd.id_path, ids, 0, ids.Length);
                string name = nd.bound_var < 0 ? "Root Problem"
                    : $"Subproblem {string.Join(".", ids)}: x{nd.bound_var + 1} {(nd.is_ceil_child != 0 ? ">=" : "<=")} {nd.bound_val}";
                if (nd.outcome == 7) { Log($"{name}: Maximum recursion depth reached → prune."); return; }
                Log($"{name}: Constraints: {Constraints(recs.Count - 1)}");
                Log($"{name}: Solving LP relaxation with {(nd.algo == 1 ? "Dual Simplex" : "Primal Simplex")}...");
                LpText(ref nd, pivots, x);
                if (nd.outcome == 0) { Log($"{name}: LP relaxation infeasible or error: {LpxNative.StatusMessage(nd.lp_status)}"); return; }
                if (nd.outcome == 1) { Log($"{name}: Invalid Simplex result (missing Solution, Tableau, Basis, or VarNames)."); return; }
                Log($"{name} LP solution: z* = {F3(nd.z)}, x* = [{xs}]");
                if (nd.outcome == 2) { Log($"{name}: Solution x* = [{xs}] is infeasible for constraints."); return; }
                if (nd.outcome == 3) { Log($"{name}: Pruned by bound (z* ≤ current best {F3(best)})."); return; }
                if (nd.outcome == 4) { best = nd.z; Log($"{name} is integer feasible. Updated BestObjective = {F3(best)}"); return; }
                for (int i = 0; i < n; i++)
                {
                    double frac = x[i] - Math.Floor(x[i]);
                    if (frac > EPS && (1 - frac) > EPS)
                        Log($"Checking x{i + 1} = {F6(x[i])}, fracPart = {F6(frac)}, distance to 0.5 = {F6(Math.Abs(frac - 0.5))}");
                }
                if (nd.outcome == 6) { Log($"{name}: No fractional variable found but solution not integral → prune."); return; }
                string xn = $"x{nd.branch_var + 1}";
                Log($"{name}: Branching on {xn} = {F3(x[nd.branch_var])} (floor={nd.floor_val}, ceil={nd.ceil_val})");
                string prefix = ids.Length == 0 ? "" : string.Join(".", ids) + ".";
                Log($"{name}: → Subproblem {prefix}{counter}: {xn} >= {nd.ceil_val} (ceil first)");
                Log($"{name}: → Subproblem {prefix}{counter + 1}: {xn} <= {nd.floor_val}");
                counter += 2;
            };

            var opt = new LpxOptions();
            LpxNative.lpx_default_options(ref opt);
            var bestX = new double[n];
            int rc = LpxNative.lpx_bnb_simplex(m, n, (int)problem.ObjectiveSense, A, rel, b, problem.C, ref opt,
                1 /* LPX_BNB_WANT_HISTORY: the root tableau is part of the result */, out int found, out double bestZ,
                bestX, out int nNodes, out long lpPivots, out int rootStatus, cb, IntPtr.Zero);
            GC.KeepAlive(cb);
            if (rc != 0) throw new Exception(LpxNative.LastError());
            if (rootStatus < 0) return new SimplexResult { Report = "LP relaxation infeasible", Summary = "Error: Infeasible" };
            if (dualRoot) return new SimplexResult { Report = "Invalid Simplex result", Summary = "Error: Invalid result" };

            var sbr = new StringBuilder();
            sbr.AppendLine("Branch & Bound Finished.");
            if (found == 0) sbr.AppendLine("No integer-feasible solution found.");
            else
            {
                sbr.AppendLine($"Best integer z* = {F3(bestZ)}");
                sbr.AppendLine($"Best integer x* = [{string.Join(", ", bestX.Select(F3))}]");
            }
            int mm = rootTableau != null ? rootTableau.GetLength(0) - 1 : m;
            return new SimplexResult
            {
                Report = sbr.ToString(), Summary = sbr.ToString(), OptimalValue = found != 0 ? bestZ : double.NegativeInfinity,
                Solution = found != 0 ? bestX : null, Tableau = rootTableau, Basis = rootBasis,
                VarNames = Enumerable.Range(0, n).Select(j => $"x{j + 1}").Concat(Enumerable.Range(0, mm).Select(j => $"c{j + 1}")).ToArray()
            };
        }
    }
}

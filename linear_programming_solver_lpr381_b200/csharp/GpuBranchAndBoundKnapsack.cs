// Drop-in ILPAlgorithm for BranchAndBoundKnapsack: same contract as BranchAndBoundKnapsack.Solve
// (Models/BranchAndBoundKnapsack.cs:58-407).  The whole best-first search runs on the GPU
// (lpx_bnb_knapsack: heap, ComputeRelaxation of both children, incumbent updates in the reference's
// order); the engine hands back one record per expanded or closed node, in pop order, and the log of
// :127-327 is re-emitted from them.  The C# twin of BranchAndBoundKnapsack::Solve in
// ../host/bnb_host.cpp (compiled and tested here, byte-equal to the restated reference text); shipped
// as source because this image has no .NET toolchain.
using System;
using System.Collections.Generic;
using System.Linq;
using System.Runtime.InteropServices;
using System.Text;

namespace Linear_Programming_Solver.Models
{
    public class GpuBranchAndBoundKnapsack : ILPAlgorithm
    {
        private const double EPS = 1e-9;
        private StringBuilder _report;
        private Action<string, bool[,]> _log;
        private int _flushPos;

        private void FlushDelta()                                   // :37-48
        {
            if (_log == null) return;
            string full = _report.ToString();
            if (full.Length > _flushPos)
            {
                string delta = full.Substring(_flushPos);
                _flushPos = full.Length;
                _log(delta, null);
            }
        }

        public SimplexResult Solve(LPProblem problem, Action<string, bool[,]> updatePivot = null)
        {
            _log = updatePivot ?? ((s, h) => { });
            _report = new StringBuilder();
            _flushPos = 0;
            if (problem == null) throw new ArgumentNullException(nameof(problem));
            if (problem.Constraints == null || problem.Constraints.Count != 1)
                throw new Exception("Knapsack solver requires exactly one constraint (weights and capacity).");
            var cons = problem.Constraints[0];
            if (cons.Relation != Rel.LE) throw new Exception("Knapsack solver requires a <= constraint.");
            int n = problem.NumVars;
            double capacity = cons.B;
            var profit = problem.C;
            var weight = new double[n];
            Array.Copy(cons.A, weight, n);                           // throws on a short row, like cons.A[i] upstream

            // ratio ordering for the header (:75-94); the engine returns the same ordering (rank_order)
            double Ratio(int i) => weight[i] > 0 ? profit[i] / weight[i] : double.PositiveInfinity;
            var order = Enumerable.Range(0, n).OrderByDescending(Ratio).ThenByDescending(i => profit[i]).ToArray();
            _report.AppendLine("Branch and Bound Knapsack Algorithm");
            _report.AppendLine("=================================================");
            _report.AppendLine("Ratio Test:");
            _report.AppendLine("Item\tci/ai\tRank");
            for (int s = 0; s < n; s++) _report.AppendLine($"{order[s] + 1}\t{Ratio(order[s]):0.###}\t{s + 1}");
            _report.AppendLine();
            _report.AppendLine("-------------------------------------------------");

            // relaxed[] of ComputeRelaxation (:431-491) from the assignment and where the greedy pass stopped
            double[] Relaxed(LpxKnapEval e)
            {
                var r = new double[n];
                var a = new byte[n];
                Marshal.Copy(e.assigned, a, 0, n);
                double fixedW = 0.0;
                for (int i = 0; i < n; i++)
                    if ((sbyte)a[i] == 1) { r[i] = 1.0; fixedW += weight[i]; }
                if (fixedW > capacity + EPS) return r;               // early return (:455-456)
                for (int s = 0; s < e.break_rank && s < n; s++)
                    if ((sbyte)a[order[s]] < 0) r[order[s]] = 1.0;
                if (e.frac_rank >= 0) r[order[e.frac_rank]] = e.frac;
                return r;
            }
            void VectorLines(LpxKnapEval e)
            {
                var r = Relaxed(e);
                int fracOrig = e.frac_rank >= 0 ? order[e.frac_rank] : -1;
                for (int i = 0; i < n; i++)
                    _report.AppendLine($"{(i == fracOrig ? ">" : " ")}\tx{i + 1}\t=\t{r[i]:0.###}");
            }

            LpxKnapPopFn onPop = (ref LpxKnapPop p, IntPtr left, IntPtr right, IntPtr user) =>
            {
                var lab = new int[p.label_len];
                Marshal.Copy(p.label, lab, 0, p.label_len);
                string label = string.Join(".", lab);
                _report.AppendLine(label == "0" ? "Sub-Problem 0" : $"Sub-Problem {label}");
                _report.AppendLine();
                VectorLines(p.relax);
                _report.AppendLine();
                if (p.closed != 0)                                    // :147-177
                {
                    if (p.closed == 3) _report.AppendLine("\tINFEASIBLE");
                    else
                    {
                        _report.AppendLine($"\tz = {Math.Round(p.relax.bound, 6):0.###}");
                        _report.AppendLine(p.closed == 1 ? "\tBEST CANDIDATE" : "\tCANDIDATE");
                    }
                    _report.AppendLine("------------------------------------------------");
                    FlushDelta();
                    return;
                }
                _report.AppendLine("------------------------------------------------");
                _report.AppendLine();
                FlushDelta();
                for (int side = 0; side < 2; side++)                  // :207-327
                {
                    var e = Marshal.PtrToStructure<LpxKnapEval>(side == 0 ? left : right);
                    string cl = label == "0" ? (side + 1).ToString() : $"{label}.{side + 1}";
                    _report.AppendLine($"-- Node {cl} branching (x{e.var + 1}={side}):");
                    VectorLines(e);
                    _report.AppendLine($"\tBound = {e.bound:0.###}, Capacity = {e.weight:0.###}");
                    if (e.decision == 1) _report.AppendLine(side == 0 ? "\tINFEASIBLE" : "\tINFEASIBLE ");
                    else if (e.decision == 2 || e.decision == 4) _report.AppendLine($"\tCANDIDATE {cl}");
                    _report.AppendLine("------------------------------------------------");
                    _report.AppendLine();
                    FlushDelta();
                }
            };

            var opt = new LpxOptions();
            LpxNative.lpx_default_options(ref opt);
            var bestX = new int[n];
            var engineOrder = new int[n];
            int rc = LpxNative.lpx_bnb_knapsack(n, profit, weight, capacity, ref opt, out int found, out double best,
                bestX, out long evals, out long pops, engineOrder, onPop, IntPtr.Zero);
            GC.KeepAlive(onPop);
            if (rc != 0) throw new Exception(LpxNative.LastError());
            if (!engineOrder.SequenceEqual(order)) throw new Exception("internal: ratio ordering mismatch between host and engine");

            _report.AppendLine();                                     // :333-368
            _report.AppendLine("Final Report:");
            _report.AppendLine("Branch & Bound Knapsack Finished.");
            _report.AppendLine();
            if (found == 0) _report.AppendLine("Status: NO FEASIBLE CANDIDATE");
            else
            {
                _report.AppendLine("Status: BEST CANDIDATE FOUND");
                for (int j = 0; j < n; j++) _report.AppendLine($"  x{j + 1} = {bestX[j]}");
                _report.AppendLine($"  z* = {Math.Round(best, 6):0.###}");
            }
            _report.AppendLine();
            _report.AppendLine();
            _report.AppendLine("Summary:");
            if (found == 0) _report.AppendLine("No feasible candidate found.");
            else
            {
                _report.AppendLine($"Best Candidate = {Math.Round(best, 6):0.###}");
                _report.AppendLine("Best x* = [" + string.Join(", ", bestX) + "]");
            }
            FlushDelta();

            var fin = new StringBuilder();                            // :371-405
            fin.AppendLine("Final Report:");
            fin.AppendLine("Branch & Bound Knapsack Finished.");
            fin.AppendLine();
            if (found == 0) fin.AppendLine("Status: INFEASIBLE");
            else
            {
                fin.AppendLine("Status: BEST CANDIDATE FOUND");
                for (int i = 0; i < n; i++) fin.AppendLine($"  x{i + 1} = {bestX[i]}");
                fin.AppendLine($"  z* = {best:0.###}");
            }
            fin.AppendLine();
            fin.AppendLine("Summary:");
            if (found == 0) fin.AppendLine("No feasible candidate found.");
            else
            {
                fin.AppendLine($"Best Candidate = {best:0.###}");
                fin.AppendLine("Best x* = [" + string.Join(", ", bestX) + "]");
            }
            return new SimplexResult { Report = fin.ToString(), Summary = "" };
        }
    }
}

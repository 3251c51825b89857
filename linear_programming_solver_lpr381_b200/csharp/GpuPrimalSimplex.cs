// Drop-in ILPAlgorithm for "Primal Simplex": same contract as PrimalSimplex.Solve
// (Models/PrimalSimplex.cs:57-127), arithmetic on the GPU through liblpx.so.
// It reuses the reference's own text helpers, which must be made `internal static`
// (AppendCanonicalForm, AppendTableau, FinalizeReport, ExpandEqualitiesToInequalities) — see
// INTEGRATION.md for the five-line patch to PrimalSimplex.cs and LPSolver.cs.
using System;
using System.Linq;
using System.Text;

namespace Linear_Programming_Solver.Models
{
    internal class GpuPrimalSimplex : ILPAlgorithm
    {
        public SimplexResult Solve(LPProblem original, Action<string, bool[,]> updatePivot = null)
        {
            var model = original.Clone();
            int m = model.Constraints.Count, n = model.NumVars;
            var A = new double[m * n];
            var rel = new int[m];
            var b = new double[m];
            for (int i = 0; i < m; i++)
            {
                Array.Copy(model.Constraints[i].A, 0, A, i * n, n);   // throws like BuildTableau on short rows
                rel[i] = (int)model.Constraints[i].Relation;          // LE=0, GE=1, EQ=2
                b[i] = model.Constraints[i].B;
            }
            var opt = new LpxOptions();
            LpxNative.lpx_default_options(ref opt);
            LpxNative.lpx_tableau_dims(m, n, rel, out int rows, out int cols);
            int cap = opt.max_iterations;
            var pivots = new int[2 * cap];
            var basisOut = new int[rows - 1];
            var x = new double[n];
            var T = new double[rows * cols];
            // the reference prints every iteration; ask the engine for all of them in one call
            int histCap = updatePivot != null ? Math.Min(cap + 1, (int)(((long)1 << 28) / Math.Max(1, rows * cols))) : 0;
            var hist = histCap > 0 ? new double[(long)histCap * rows * cols] : null;
            int rc = LpxNative.lpx_primal_solve(m, n, (int)model.ObjectiveSense, A, rel, b, model.C, ref opt,
                out int status, out int nPivots, pivots, cap, basisOut, x, out double z, T, hist, histCap);
            if (rc != 0) throw new Exception(LpxNative.LastError());
            if (status < 0) throw new Exception(LpxNative.StatusMessage(status));   // same messages as upstream

            // text: the reference's own helpers on the engine's numbers
            if (model.ObjectiveSense == Sense.Min) for (int i = 0; i < n; i++) model.C[i] = -model.C[i];
            var tableauModel = PrimalSimplex.ExpandEqualitiesToInequalities(model);
            var report = new StringBuilder();
            PrimalSimplex.AppendCanonicalForm(report, tableauModel);
            int mm = rows - 1;
            var varNames = Enumerable.Range(0, n).Select(j => $"x{j + 1}")
                .Concat(Enumerable.Range(0, mm).Select(j => $"c{j + 1}")).ToArray();
            if (updatePivot != null)
            {
                var basis = Enumerable.Range(n, mm).ToArray();
                for (int k = 0; k <= Math.Min(nPivots, histCap - 1); k++)
                {
                    var Tk = new double[rows, cols];
                    Buffer.BlockCopy(hist, k * rows * cols * 8, Tk, 0, rows * cols * 8);
                    bool[,] hl = null;
                    if (k > 0)
                    {
                        int e = pivots[2 * (k - 1)], l = pivots[2 * (k - 1) + 1];
                        basis[l] = e;
                        hl = new bool[rows, cols];
                        for (int j = 0; j < cols; j++) hl[l, j] = true;
                        for (int i = 0; i < rows; i++) hl[i, e] = true;
                    }
                    var sb = new StringBuilder();
                    PrimalSimplex.AppendTableau(sb, Tk, basis, varNames, k);
                    updatePivot(sb.ToString(), hl);
                }
            }
            var Tfinal = new double[rows, cols];
            Buffer.BlockCopy(T, 0, Tfinal, 0, rows * cols * 8);
            if (status == 1) report.AppendLine("UNBOUNDED");
            return PrimalSimplex.FinalizeReport(report, Tfinal, basisOut, varNames, status == 1 ? "UNBOUNDED" : "OPTIMAL");
        }
    }
}

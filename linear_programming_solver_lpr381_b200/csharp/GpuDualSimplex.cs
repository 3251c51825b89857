// Drop-in ILPAlgorithm for "Dual Simplex": same contract as DualSimplex.Solve
// (Models/DualSimplex.cs:15-114), arithmetic on the GPU through liblpx.so (lpx_dual_solve).
// It reuses the reference's own text helpers, which must be made `internal static`
// (DualSimplex.AppendTableau, DualSimplex.FinalizeReport) — see INTEGRATION.md.  The C# twin of
// DualSimplex::Solve in ../host/simplex_host.cpp (compiled and tested here); shipped as source
// because this image has no .NET toolchain.
//
// Quirks kept on purpose: the silent ForceDualFeasibility pivots (:195-228) are not printed, "Iteration 0"
// is the tableau after them; the optimal tableau is printed once more with row 0 highlighted (:58-72);
// the result carries Report and Summary only (:310), so LPSolver.FinalTableau stays null and
// BranchAndBound rejects it as "Invalid Simplex result"; the iteration limit is the literal 10000 (:39).
using System;
using System.Linq;
using System.Text;

namespace Linear_Programming_Solver.Models
{
    internal class GpuDualSimplex : ILPAlgorithm
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
            int cap = 10000 + 128;                                    // <= 100 silent + <= 10000 dual pivots
            var pivots = new int[2 * cap];
            var basisOut = new int[rows - 1];
            var x = new double[n];
            var T = new double[rows * cols];
            int status, nPivots, silent;
            double z;
            double[] hist = null;
            int histCap = 0;
            if (updatePivot != null)
            {
                // a first pass sizes the history: the engine returns every iteration's tableau in one call
                int rc0 = LpxNative.lpx_dual_solve(m, n, (int)model.ObjectiveSense, A, rel, b, model.C, ref opt, out status,
                    out nPivots, out silent, null, 0, null, null, out z, null, null, 0);
                if (rc0 != 0) throw new Exception(LpxNative.LastError());
                histCap = nPivots - silent + 1;
                hist = new double[(long)histCap * rows * cols];
            }
            int rc = LpxNative.lpx_dual_solve(m, n, (int)model.ObjectiveSense, A, rel, b, model.C, ref opt, out status,
                out nPivots, out silent, pivots, cap, basisOut, x, out z, T, hist, histCap);
            if (rc != 0) throw new Exception(LpxNative.LastError());

            int mm = rows - 1;
            var varNames = Enumerable.Range(0, n).Select(j => $"x{j + 1}")
                .Concat(Enumerable.Range(0, mm).Select(j => $"c{j + 1}")).ToArray();
            var Tfinal = new double[rows, cols];
            Buffer.BlockCopy(T, 0, Tfinal, 0, rows * cols * 8);
            if (updatePivot != null)
            {
                // basis after the silent pivots (:24), then one chunk per iteration with the pivot cross
                var basis = Enumerable.Range(n, mm).ToArray();
                for (int k = 0; k < silent; k++) basis[pivots[2 * k + 1]] = pivots[2 * k];
                for (int k = 0; k < histCap; k++)
                {
                    var Tk = new double[rows, cols];
                    Buffer.BlockCopy(hist, k * rows * cols * 8, Tk, 0, rows * cols * 8);
                    bool[,] hl = null;
                    if (k > 0)
                    {
                        int e = pivots[2 * (silent + k - 1)], l = pivots[2 * (silent + k - 1) + 1];
                        basis[l] = e;
                        hl = new bool[rows, cols];
                        for (int j = 0; j < cols; j++) hl[l, j] = true;
                        for (int i = 0; i < rows; i++) hl[i, e] = true;
                    }
                    var sb = new StringBuilder();
                    DualSimplex.AppendTableau(sb, Tk, basis, varNames, k);
                    updatePivot(sb.ToString(), hl);
                }
                if (status == 0)
                {
                    // the optimal tableau once more, "z row" (row 0 of the mask) highlighted (:58-72)
                    var hl = new bool[rows, cols];
                    for (int j = 0; j < cols; j++) hl[0, j] = true;
                    var sb = new StringBuilder();
                    DualSimplex.AppendTableau(sb, Tfinal, basisOut, varNames, nPivots - silent + 1);
                    updatePivot(sb.ToString(), hl);
                }
            }
            if (status == -3) throw new Exception("Iteration limit exceeded (Dual Simplex).");
            var report = new StringBuilder();
            if (status == 2) report.AppendLine("INFEASIBLE (no entering column found)");
            return DualSimplex.FinalizeReport(report, Tfinal, basisOut, varNames, status == 2 ? "INFEASIBLE" : "OPTIMAL");
        }
    }
}

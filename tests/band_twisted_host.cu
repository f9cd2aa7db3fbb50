// CPU check of the twisted band LU / solve of the warp-per-system kernel: the sweeps in
// rehuel_b200/csrc/band_twisted.cuh are __host__ __device__, so the code the GPU lanes run is run
// here on the host, the two lanes of a matrix one after the other, the shuffles replaced by plain
// copies -- against a dense complex solve with partial pivoting.
//   nvcc -O1 -std=c++17 -o build/band_twisted_host tests/band_twisted_host.cu && build/band_twisted_host
// (tests/test_capi_host.py builds and runs it; test infrastructure, not a product path.)
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../rehuel_b200/csrc/band_twisted.cuh"

using cd = std::complex<double>;

static unsigned long long rng_state = 0x5EEDull;
static double uni() // splitmix64 -> (-1, 1)
{
	unsigned long long z = (rng_state += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z ^= z >> 31;
	return 2.0 * ((double)(z >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
}

static std::vector<cd> dense_solve(int n, std::vector<cd> A, std::vector<cd> b)
{
	for (int k = 0; k < n; ++k) {
		int p = k;
		for (int r = k + 1; r < n; ++r)
			if (std::abs(A[r * n + k]) > std::abs(A[p * n + k])) p = r;
		for (int c = 0; c < n; ++c) std::swap(A[k * n + c], A[p * n + c]);
		std::swap(b[k], b[p]);
		for (int r = k + 1; r < n; ++r) {
			const cd l = A[r * n + k] / A[k * n + k];
			for (int c = k; c < n; ++c) A[r * n + c] -= l * A[k * n + c];
			b[r] -= l * b[k];
		}
	}
	for (int k = n - 1; k >= 0; --k) {
		for (int c = k + 1; c < n; ++c) b[k] -= A[k * n + c] * b[c];
		b[k] /= A[k * n + k];
	}
	return b;
}

template <int N, int KB>
static int run_case(bool is_c, bool dominant, int reps, double tol)
{
	using G = rehuel::Tw<N, KB>;
	int bad = 0, flagged = 0;
	double worst = 0.0;
	for (int rep = 0; rep < reps; ++rep) {
		std::vector<cd> A(N * N, cd(0, 0)), b(N);
		std::vector<double> tr(G::HP, 0.0), ti(G::HP, 0.0), br(G::HP, 0.0), bi(G::HP, 0.0);
		for (int i = 0; i < N; ++i) {
			for (int d = -KB; d <= KB; ++d) {
				const int j = i + d;
				cd v(0, 0);
				if (j >= 0 && j < N) {
					v = cd(uni(), is_c ? uni() : 0.0);
					if (d == 0 && dominant) v += cd(2.0 * KB + 2.0, is_c ? 1.5 : 0.0);
					A[i * N + j] = v;
				}
				rehuel::tw_put<N, KB>(tr.data(), br.data(), i, d, v.real());
				rehuel::tw_put<N, KB>(ti.data(), bi.data(), i, d, v.imag());
			}
			b[i] = cd(uni(), is_c ? uni() : 0.0);
		}
		// ---- factorisation: the two sweeps, the exchange, the separator block (by either lane, own orientation)
		double s0r[KB][KB], s0i[KB][KB], s1r[KB][KB], s1i[KB][KB], o0r[KB][KB], o0i[KB][KB], o1r[KB][KB], o1i[KB][KB];
		bool f = rehuel::tw_lu_sweep<N, KB>(tr.data(), ti.data(), is_c, G::NP0, s0r, s0i);
		f |= rehuel::tw_lu_sweep<N, KB>(br.data(), bi.data(), is_c, G::NP1, s1r, s1i);
		for (int j = 0; j < KB; ++j)
			for (int c = 0; c < KB; ++c) {
				o0r[j][c] = s1r[KB - 1 - j][KB - 1 - c];
				o0i[j][c] = s1i[KB - 1 - j][KB - 1 - c];
				o1r[j][c] = s0r[KB - 1 - j][KB - 1 - c];
				o1i[j][c] = s0i[KB - 1 - j][KB - 1 - c];
			}
		double q0r[KB * KB], q0i[KB * KB] = {0}, q1r[KB * KB], q1i[KB * KB] = {0};
		f |= rehuel::tw_separator<N, KB>(tr.data(), ti.data(), is_c, G::NP0, s0r, s0i, o0r, o0i, q0r, q0i);
		f |= rehuel::tw_separator<N, KB>(br.data(), bi.data(), is_c, G::NP1, s1r, s1i, o1r, o1i, q1r, q1i);
		flagged += f ? 1 : 0;
		// ---- two solves with the same factors (the Newton iteration reuses them)
		for (int pass = 0; pass < 2; ++pass) {
			std::vector<double> xr(N), xi(N);
			for (int i = 0; i < N; ++i) {
				if (pass) b[i] = cd(uni(), is_c ? uni() : 0.0);
				xr[i] = b[i].real();
				xi[i] = b[i].imag();
			}
			double g0r[KB], g0i[KB], g1r[KB], g1i[KB], h0r[KB], h0i[KB], h1r[KB], h1i[KB], x0r[KB], x0i[KB], x1r[KB], x1i[KB];
			rehuel::tw_forward<N, KB>(tr.data(), ti.data(), is_c, G::NP0, xr.data(), xi.data(), 1, g0r, g0i);
			rehuel::tw_forward<N, KB>(br.data(), bi.data(), is_c, G::NP1, xr.data() + N - 1, xi.data() + N - 1, -1, g1r, g1i);
			for (int j = 0; j < KB; ++j) {
				h0r[j] = g1r[KB - 1 - j];
				h0i[j] = g1i[KB - 1 - j];
				h1r[j] = g0r[KB - 1 - j];
				h1i[j] = g0i[KB - 1 - j];
			}
			rehuel::tw_separator_solve<N, KB>(q0r, q0i, is_c, G::NP0, xr.data(), xi.data(), 1, g0r, g0i, h0r, h0i, x0r, x0i);
			rehuel::tw_separator_solve<N, KB>(q1r, q1i, is_c, G::NP1, xr.data() + N - 1, xi.data() + N - 1, -1, g1r, g1i, h1r, h1i, x1r,
			                                  x1i);
			rehuel::tw_backward<N, KB>(tr.data(), ti.data(), is_c, G::NP0, xr.data(), xi.data(), 1, x0r, x0i);
			rehuel::tw_backward<N, KB>(br.data(), bi.data(), is_c, G::NP1, xr.data() + N - 1, xi.data() + N - 1, -1, x1r, x1i);
			for (int j = 0; j < KB; ++j) {
				xr[G::NP0 + j] = x0r[j];
				xi[G::NP0 + j] = x0i[j];
			}
			if (!dominant) continue; // without dominance only the flag is of interest
			const std::vector<cd> want = dense_solve(N, A, b);
			double scale = 0.0, err = 0.0;
			for (int i = 0; i < N; ++i) scale = std::max(scale, std::abs(want[i]));
			for (int i = 0; i < N; ++i) err = std::max(err, std::abs(cd(xr[i], xi[i]) - want[i]) / scale);
			for (int j = 0; j < KB; ++j) // both lanes must hold the same separator solution
				err = std::max(err, std::abs(cd(x1r[j], x1i[j]) - cd(x0r[KB - 1 - j], x0i[KB - 1 - j])) / scale);
			worst = std::max(worst, err);
			if (!(err <= tol)) ++bad;
		}
	}
	std::printf("N=%2d KB=%d %s %s: worst rel. error %.2e, would-pivot flags %d/%d\n", N, KB, is_c ? "complex" : "real   ",
	            dominant ? "dominant" : "random  ", worst, flagged, reps);
	if (dominant && flagged) ++bad;          // a dominant matrix never needs an interchange
	if (!dominant && flagged * 10 < reps * 9 && N >= 32) ++bad; // a random 64-row band matrix practically always does
	return bad;
}

int main()
{
	int bad = 0;
	for (int c = 0; c < 2; ++c) {
		bad += run_case<64, 2>(c, true, 50, 1e-13);
		bad += run_case<64, 2>(c, false, 50, 0.0);
		bad += run_case<8, 2>(c, true, 50, 1e-13);
		bad += run_case<11, 3>(c, true, 50, 1e-13);
		bad += run_case<9, 1>(c, true, 50, 1e-13);
		bad += run_case<64, 4>(c, true, 50, 1e-13);
		bad += run_case<14, 4>(c, true, 50, 1e-13);
		bad += run_case<33, 2>(c, true, 50, 1e-13);
	}
	std::printf(bad ? "FAILED\n" : "OK\n");
	return bad ? 1 : 0;
}

/* The reference reserves C_interface_test/main.c (an empty file) for a C binding; this is that
 * program against include/rehuel_b200.h: a 4096-member swept Robertson ensemble, RADAU_IIA_53.
 *   cc -std=c11 -Iinclude examples/c_interface_test.c -Lrehuel_b200 -lrehuel_b200 -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "rehuel_b200.h"

int main(void)
{
	const size_t M = 4096;
	double *params = malloc(sizeof(double) * 3 * M);
	const double y0[3] = {1.0, 0.0, 0.0};
	for (size_t i = 0; i < M; ++i) { /* a simple deterministic sweep of the three rate constants */
		double u = (double)i / (double)(M - 1) - 0.5;
		params[0 * M + i] = 0.04 * pow(10.0, u);
		params[1 * M + i] = 1e4 * pow(10.0, -u);
		params[2 * M + i] = 3e7 * pow(10.0, 0.5 * u);
	}
	rehuel_options o;
	rehuel_default_options(&o);
	o.rel_tol = o.abs_tol = 1e-6;
	o.newton_tol = 0.1 * 1e-6; /* examples/example_equations.cpp:220 */
	rehuel_result r;
	int method = rehuel_name_to_method("RADAU_IIA_53");
	int problem = rehuel_name_to_problem("robertson");
	int rc = rehuel_irk_ensemble(problem, method, 0.0, 40.0, 1e-6, M, 0, y0, 1, params, &o, 0, 1, &r);
	if (rc != REHUEL_OK) {
		fprintf(stderr, "rehuel_irk_ensemble failed (%d): %s\n", rc, rehuel_last_error());
		return 2;
	}
	double worst = 0.0;
	for (size_t i = 0; i < M; ++i) {
		double mass = r.y[0 * M + i] + r.y[1 * M + i] + r.y[2 * M + i];
		if (fabs(mass - 1.0) > worst) worst = fabs(mass - 1.0);
	}
	printf("solved %zu systems with %s in %.3f ms (kernel %.3f ms); accepted %.1f%% of %llu attempts; "
	       "failed members %llu; max |y1+y2+y3-1| = %.2e; y(member 0) = %.12g %.12g %.12g\n",
	       M, rehuel_method_to_name(method), r.elapsed_ms, r.kernel_ms, 100.0 * r.accept_frac, r.sum.attempt,
	       r.sum.n_failed, worst, r.y[0], r.y[M], r.y[2 * M]);
	int ok = r.sum.n_failed == 0 && worst < 1e-12;
	rehuel_result_free(&r);
	free(params);
	return ok ? 0 : 1;
}

/* Plain-C caller of libnode4j_b200.so — what a JNI / Panama shim does, with no Python, torch or CUDA headers involved.
 *
 * Runs the reference's own known-answer test through the C ABI (include/node4j.h): MultiStepAdjointTest.java:174-232 —
 * OdeVertex(Dense 2->2, identity) with InputStep(DormandPrince54(1e-12, 1e-6, 1e-20, 1e2), 1, interpolate=true,
 * needTimeGradient=true), f(y) = y*A + b, A = 0.2*ones(2,2), b = 3, t = linspace(1, 8, 10), y0 = linspace(-1.23, 2.34, 4),
 * lossgrad = linspace(3, 13, 40) with the test's sign flips (:201-204) — and compares with the golden numbers the pytest
 * wrapper passes in a text file (tests/golden/reference_vectors.json: the torchdiffeq vectors quoted in that Java test),
 * with the Java test's tolerances.  Without a CUDA device it checks the documented failure instead: node4j_create returns
 * N4J_ERR_CUDA with a message, never a CPU fallback.
 *
 *   gcc -O1 -I include tests/abi_driver.c -o abi_driver -L neuralode4j_b200 -lnode4j_b200 -Wl,-rpath,$PWD/neuralode4j_b200
 *   ./abi_driver golden.txt      (40 ys, 4 dL/dy0, 6 dL/dtheta, 10 dL/dt)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "node4j.h"

#define CHECK(call)                                                                       \
    do {                                                                                  \
        int32_t rc_ = (call);                                                             \
        if (rc_ != N4J_OK) {                                                              \
            fprintf(stderr, "%s -> %d: %s\n", #call, (int)rc_, node4j_last_error());      \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

static int read_doubles(FILE* f, double* dst, int n) {
    for (int i = 0; i < n; ++i)
        if (fscanf(f, "%lf", &dst[i]) != 1) return 0;
    return 1;
}

int main(int argc, char** argv) {
    if (node4j_abi_version() != N4J_ABI_VERSION) { fprintf(stderr, "ABI version %d\n", (int)node4j_abi_version()); return 1; }

    n4j_layer_desc layer;
    memset(&layer, 0, sizeof(layer));
    layer.kind = N4J_LAYER_DENSE; layer.activation = N4J_ACT_IDENTITY; layer.n_in = 2; layer.n_out = 2; layer.has_bias = 1; layer.eps = 1e-5f;
    n4j_graph_desc graph;
    memset(&graph, 0, sizeof(graph));
    graph.n_layers = 1; graph.rank = 2; graph.shape[0] = 2; graph.shape[1] = 2; graph.layers = &layer;
    n4j_solver_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abs_tol = 1e-12; cfg.rel_tol = 1e-6; cfg.min_step = 1e-20; cfg.max_step = 1e2;      /* MultiStepAdjointTest.java:316-317 */
    cfg.safety = 0.9; cfg.max_growth = 10.0; cfg.min_reduction = 0.2; cfg.order = 5;        /* AdaptiveRungeKuttaStepPolicy.java:43-45 */

    n4j_handle* h = NULL;
    if (node4j_device_count() < 1) {
        const int32_t rc = node4j_create(&graph, &cfg, 0, &h);
        if (rc != N4J_ERR_CUDA || h != NULL || strlen(node4j_last_error()) == 0) {
            fprintf(stderr, "expected N4J_ERR_CUDA without a device, got %d\n", (int)rc);
            return 1;
        }
        printf("ABI DRIVER: NO DEVICE -> N4J_ERR_CUDA (%s)\n", node4j_last_error());
        return 0;
    }
    if (argc < 2) { fprintf(stderr, "usage: abi_driver golden.txt\n"); return 2; }
    FILE* f = fopen(argv[1], "r");
    if (!f) { perror(argv[1]); return 2; }
    double ys_ref[40], dy0_ref[4], dth_ref[6], dt_ref[10];
    if (!read_doubles(f, ys_ref, 40) || !read_doubles(f, dy0_ref, 4) || !read_doubles(f, dth_ref, 6) || !read_doubles(f, dt_ref, 10)) {
        fprintf(stderr, "short golden file\n");
        return 2;
    }
    fclose(f);

    CHECK(node4j_create(&graph, &cfg, 0, &h));
    if (node4j_num_params(h) != 6 || node4j_num_real_params(h) != 6 || node4j_state_len(h) != 4) { fprintf(stderr, "parameter layout\n"); return 1; }
    /* flat DL4J parameters: W f-order [nIn,nOut] then bias (MultiStepAdjointTest.java:216) */
    float params[6] = {0.2f, 0.2f, 0.2f, 0.2f, 3.f, 3.f};
    CHECK(node4j_bind_params(h, params, 6));
    float y0[4], t[10], ys[40], lossgrad[40];
    for (int i = 0; i < 4; ++i) y0[i] = (float)(-1.23 + (2.34 + 1.23) * i / 3.0);
    for (int i = 0; i < 10; ++i) t[i] = (float)(1.0 + 7.0 * i / 9.0);
    n4j_stats st;
    CHECK(node4j_forward(h, y0, t, 10, 1, ys, &st));                    /* out is [B=2, F=2, T=10] */
    int bad = 0;
    for (int i = 0; i < 40; ++i)
        if (fabs(ys[i] - ys_ref[i]) > 1e-3) { fprintf(stderr, "ys[%d] = %g, golden %g\n", i, ys[i], ys_ref[i]); ++bad; }
    const long long nfe_fwd = (long long)st.nfe, acc_fwd = (long long)st.accepted;
    for (int i = 0; i < 40; ++i) lossgrad[i] = (float)(3.0 + 10.0 * i / 39.0);
    for (int j = 0; j < 2; ++j) lossgrad[(0 * 2 + j) * 10 + 1] *= -1.f;  /* lossgrad[0, :, 1].negi()  (:201-204, size(0)/2 == 1) */
    float grad[6] = {0, 0, 0, 0, 0, 0}, dy0[4], dt[10];
    CHECK(node4j_backward(h, NULL, lossgrad, t, 10, N4J_TIMEGRAD_CALC, grad, dy0, dt, &st));
    for (int i = 0; i < 4; ++i)
        if (fabs(dy0[i] - dy0_ref[i]) > 1e-3) { fprintf(stderr, "dL/dy0[%d] = %g, golden %g\n", i, dy0[i], dy0_ref[i]); ++bad; }
    for (int i = 0; i < 6; ++i)
        if (fabs(grad[i] / dth_ref[i] - 1.0) > 1e-4) { fprintf(stderr, "dL/dtheta[%d] = %g, golden %g\n", i, grad[i], dth_ref[i]); ++bad; }
    for (int i = 0; i < 10; ++i)
        if (fabs(dt[i] / dt_ref[i] - 1.0) > 1e-4) { fprintf(stderr, "dL/dt[%d] = %g, golden %g\n", i, dt[i], dt_ref[i]); ++bad; }
    /* error behaviour: time must have at least two entries (AdaptiveRungeKuttaSolver.java:108-110 -> IllegalArgumentException) */
    if (node4j_forward(h, y0, t, 1, 1, ys, &st) != N4J_ERR_INVALID_ARGUMENT) { fprintf(stderr, "nt = 1 was accepted\n"); ++bad; }
    n4j_step_rec rec[8];
    const long long ntrial = (long long)node4j_get_step_log(h, rec, 8);
    CHECK(node4j_destroy(h));
    if (bad) return 1;
    printf("ABI DRIVER OK: forward nfe %lld accepted %lld, backward nfe %lld accepted %lld solves %lld, %lld trial records\n", nfe_fwd, acc_fwd,
           (long long)st.nfe, (long long)st.accepted, (long long)st.solves, ntrial);
    return 0;
}

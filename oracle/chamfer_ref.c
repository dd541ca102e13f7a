/* ORACLE -- test infrastructure, not product code.
 *
 * CPU restatement of the Chamfer nearest-neighbour forward of the reference's
 * extensions/chamfer_dist CUDA extension (/root/reference/README.md:62-65 names the extension and
 * how it is built; its source is on the upstream Stereo2Point branch and NOT on disk).
 * PARITY UNPINNED: there is no reference source, test or golden vector to pin this against; the
 * arithmetic below is this repo's [SPEC] (SURVEY.md 8(a) row C, section 7 "bit-exact Chamfer"):
 *   d(i,j) = ((x1-x2)^2 + (y1-y2)^2) + (z1-z2)^2   in IEEE fp32, round-to-nearest, NO fused
 *   multiply-add (compile with -ffp-contract=off), scan j ascending, update on strict '<',
 *   so ties resolve to the lowest index.  best starts at +inf / index INT32_MAX.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>

static void nn_dir(const float* q, const float* r, float* dist, int32_t* idx, int nq, int nr) {
  for (int i = 0; i < nq; ++i) {
    const float qx = q[i * 3 + 0], qy = q[i * 3 + 1], qz = q[i * 3 + 2];
    float best = INFINITY;
    int32_t bi = INT32_MAX;
    for (int j = 0; j < nr; ++j) {
      const float dx = qx - r[j * 3 + 0], dy = qy - r[j * 3 + 1], dz = qz - r[j * 3 + 2];
      const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
      const float s = xx + yy;
      const float d = s + zz;
      if (d < best) { best = d; bi = j; }
    }
    dist[i] = best;
    idx[i] = bi == INT32_MAX ? 0 : bi;   /* nothing compared less (all NaN / +inf): a valid index */
  }
}

typedef struct {
  const float *xyz1, *xyz2;
  float *dist1, *dist2;
  int32_t *idx1, *idx2;
  int B, N, M, tid, nthreads;
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  /* work item = (batch, direction); items are independent, static round-robin over threads */
  for (int it = j->tid; it < 2 * j->B; it += j->nthreads) {
    const int b = it >> 1;
    if ((it & 1) == 0)
      nn_dir(j->xyz1 + (long)b * j->N * 3, j->xyz2 + (long)b * j->M * 3, j->dist1 + (long)b * j->N,
             j->idx1 + (long)b * j->N, j->N, j->M);
    else
      nn_dir(j->xyz2 + (long)b * j->M * 3, j->xyz1 + (long)b * j->N * 3, j->dist2 + (long)b * j->M,
             j->idx2 + (long)b * j->M, j->M, j->N);
  }
  return 0;
}

/* xyz1 [B,N,3], xyz2 [B,M,3] -> dist1/idx1 [B,N], dist2/idx2 [B,M]; nthreads >= 1 host threads */
void chamfer_ref(const float* xyz1, const float* xyz2, float* dist1, int32_t* idx1, float* dist2,
                 int32_t* idx2, int B, int N, int M, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  job_t jobs[256];
  for (int t = 0; t < nthreads; ++t) {
    job_t j = {xyz1, xyz2, dist1, dist2, idx1, idx2, B, N, M, t, nthreads};
    jobs[t] = j;
    if (t > 0) pthread_create(&th[t], 0, worker, &jobs[t]);
  }
  worker(&jobs[0]);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], 0);
}

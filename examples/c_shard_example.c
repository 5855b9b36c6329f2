/* c_shard_example.c -- the multi-GPU partitioning entries of libb200osd.so from plain C; host only (needs no device).
 * A "table" of the edge midpoints of a ring of 64 control vertices is cut for 4 ranks in two ways:
 *   b200osd_shard_plan           contiguous row ranges of equal cost (every rank may then need every control vertex),
 *   b200osd_shard_plan_locality  rows dealt out by the control vertices they reference, and
 *   b200osd_shard_control_runs   the index runs of control vertices a rank has to receive per frame.
 *
 *   gcc -std=c99 -I include examples/c_shard_example.c -L opensubdiv_b200 -lb200osd -Wl,-rpath,$PWD/opensubdiv_b200 -o /tmp/shard_ex
 */
#include <stdio.h>

#include "b200osd_capi.h"

#define NV 64
#define WORLD 4

int main(void)
{
    int sizes[NV], offsets[NV], indices[2 * NV];
    int ranges[2 * WORLD], order[NV], cuts[2 * WORLD], ctrl[2 * WORLD];
    int i, r, ok = 1;

    /* row i = midpoint of edge (p(i), p(i) + 1 mod NV), listed in a scrambled order like a real table's blocks */
    for (i = 0; i < NV; ++i) {
        const int p = (i * 29) % NV;
        sizes[i] = 2;
        offsets[i] = 2 * i;
        indices[2 * i] = p;
        indices[2 * i + 1] = (p + 1) % NV;
    }
    if (b200osd_shard_plan(NV, sizes, WORLD, 1, ranges) != B200OSD_OK) { printf("shard_plan: %s\n", b200osd_last_error()); return 1; }
    if (b200osd_shard_plan_locality(NV, sizes, offsets, indices, WORLD, order, cuts, ctrl) != B200OSD_OK) {
        printf("shard_plan_locality: %s\n", b200osd_last_error());
        return 1;
    }
    for (r = 0; r < WORLD; ++r) {
        int lsz[NV], loff[NV], lidx[2 * NV], runs[2 * 4];
        int n = 0, k, q, need = 0;
        for (q = cuts[2 * r]; q < cuts[2 * r + 1]; ++q, ++n) {          /* this rank's local table */
            lsz[n] = 2;
            loff[n] = 2 * n;
            lidx[2 * n] = indices[offsets[order[q]]];
            lidx[2 * n + 1] = indices[offsets[order[q]] + 1];
        }
        k = b200osd_shard_control_runs(n, lsz, loff, lidx, 1, 4, runs);
        if (k < 0) { printf("shard_control_runs: %s\n", b200osd_last_error()); return 1; }
        printf("rank %d: contiguous rows [%d,%d) | by locality %d rows, control vertices [%d,%d) in %d run(s):", r, ranges[2 * r],
               ranges[2 * r + 1], n, ctrl[2 * r], ctrl[2 * r + 1], k);
        for (q = 0; q < k; ++q) { printf(" [%d,%d)", runs[2 * q], runs[2 * q + 1]); need += runs[2 * q + 1] - runs[2 * q]; }
        printf(" = %d of %d\n", need, NV);
        ok = ok && n == NV / WORLD && need <= NV / WORLD + 2 && k >= 1;  /* a quarter of the ring + the shared end points */
    }
    return ok ? 0 : 1;
}

/* c_abi_example.c -- the smallest client of libb200osd.so, in plain C: one stencil table (the 4 edge midpoints and the
 * centre of a quad), EvalStencils of xyz + its trivial "derivative" stream, read back.
 *
 *   gcc -std=c99 -I include examples/c_abi_example.c -L opensubdiv_b200 -lb200osd -Wl,-rpath,$PWD/opensubdiv_b200 -o /tmp/ex
 *
 * What an Osd client does through the C++ classes (include/b200osd/) is exactly these calls. */
#include <stdio.h>

#include "b200osd_capi.h"

int main(void)
{
    /* Far::StencilTable layout: sizes / offsets / indices / weights (far/stencilTable.h:107-113) */
    const int   sizes[5]    = { 2, 2, 2, 2, 4 };
    const int   offsets[5]  = { 0, 2, 4, 6, 8 };
    const int   indices[12] = { 0, 1,  1, 2,  2, 3,  3, 0,  0, 1, 2, 3 };
    const float weights[12] = { .5f, .5f, .5f, .5f, .5f, .5f, .5f, .5f, .25f, .25f, .25f, .25f };
    const float cage[4][3]  = { { 0, 0, 0 }, { 1, 0, 0 }, { 1, 1, 0.5f }, { 0, 1, 0 } };
    float refined[5][3];
    int srcDesc[3] = { 0, 3, 3 };                    /* Osd::BufferDescriptor(offset, length, stride) */
    int dstDesc[1][3] = { { 4 * 3, 3, 3 } };         /* refined points follow the 4 control points (osd/mesh.h:505-519) */
    float *dsts[1];
    int i, rc;

    b200osd_vertex_buffer *vb = b200osd_vertex_buffer_create(3, 4 + 5);
    b200osd_stencil_table *st;
    if (!vb) { printf("no CUDA device: %s\n", b200osd_last_error()); return 2; }   /* no CPU fallback */
    st = b200osd_stencil_table_create(5, 4, sizes, offsets, indices, weights, NULL, NULL, NULL, NULL, NULL, 0);
    if (!st) { printf("table: %s\n", b200osd_last_error()); return 1; }

    b200osd_vertex_buffer_update(vb, &cage[0][0], 0, 4, NULL);
    dsts[0] = b200osd_vertex_buffer_bind(vb);
    rc = b200osd_stencil_table_eval(st, b200osd_vertex_buffer_bind(vb), srcDesc, 1, dsts, (const int (*)[3])dstDesc, 0, 5, NULL);
    if (rc != B200OSD_OK) { printf("eval: %s\n", b200osd_last_error()); return 1; }
    b200osd_vertex_buffer_read(vb, &refined[0][0], 4, 5, NULL);
    b200osd_synchronize(NULL);
    for (i = 0; i < 5; ++i) printf("refined[%d] = %g %g %g\n", i, refined[i][0], refined[i][1], refined[i][2]);

    b200osd_stencil_table_destroy(st);
    b200osd_vertex_buffer_destroy(vb);
    return (refined[4][0] == 0.5f && refined[4][1] == 0.5f && refined[4][2] == 0.125f) ? 0 : 1;
}

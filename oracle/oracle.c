/*
 * oracle.c -- CPU restatement (plain C) of the reference algorithms on the DetZero
 * point-cloud hot path.  TEST INFRASTRUCTURE ONLY: nothing under detzero_b200/ may
 * import, link or call this file.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * Parity status: "parity unpinned" for the voxelizer (the arithmetic lives in
 * spconv 2.x `Point2VoxelCPU3d`, a third-party wheel that is neither vendored in
 * /root/reference nor installable here; SURVEY.md Appendix A.1 restates its published
 * semantics and the reference's call site data_processor.py:61-91 fixes how it is
 * driven).  The rotated-IoU / NMS part follows code that IS in the reference
 * (utils/detzero_utils/ops/iou3d_nms/src/iou3d_nms_kernel.cu:42-335,386-430 and
 * iou3d_nms.cpp:114-160); it is pinned against oracle/_ref (the reference's own
 * iou3d_cpu.cpp compiled in place) when that has been built.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction so the fp32
 * subtract/divide/floor sequence is IEEE-exact, Appendix A.1).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * Hard voxelization.  Follows spconv 2.x Point2VoxelCPU3d.point_to_voxel as restated in
 * SURVEY.md Appendix A.1, driven like data_processor.py:70-83 (vsize_xyz, coors_range_xyz,
 * max_num_points_per_voxel, max_num_voxels).  Internal order is zyx.
 *   points  (n, c) f32 row-major, columns x,y,z,...
 *   voxels  (max_voxels, max_pts, c) f32 -- zero-filled here
 *   coords  (max_voxels, 3) i32 [z,y,x]
 *   num_per (max_voxels) i32
 *   lookup  dense grid (gz*gy*gx) i32 workspace, must be all -1 on entry; restored on exit
 * returns number of voxels.
 * ------------------------------------------------------------------------------------ */
int oracle_points_to_voxel(const float *points, int n, int c,
                           const float *vsize_xyz, const float *range_xyz,
                           int max_pts, int max_voxels,
                           float *voxels, int32_t *coords, int32_t *num_per,
                           int32_t *lookup)
{
    /* zyx views of the xyz parameters (Appendix A.1: constructor reverses xyz -> zyx) */
    float vs[3], lo[3], hi[3];
    int grid[3];
    for (int d = 0; d < 3; ++d) {
        vs[d] = vsize_xyz[2 - d];
        lo[d] = range_xyz[2 - d];
        hi[d] = range_xyz[5 - d];
        grid[d] = (int)lroundf((hi[d] - lo[d]) / vs[d]);
    }
    memset(voxels, 0, (size_t)max_voxels * max_pts * c * sizeof(float));
    memset(num_per, 0, (size_t)max_voxels * sizeof(int32_t));
    int num_voxels = 0;
    for (int i = 0; i < n; ++i) {
        int cz[3];
        int ok = 1;
        for (int d = 0; d < 3; ++d) {
            /* float subtract, float divide, floor -- in that order, no reciprocal */
            volatile float diff = points[(size_t)i * c + (2 - d)] - lo[d];
            volatile float q = diff / vs[d];
            int ci = (int)floorf(q);
            if (ci < 0 || ci >= grid[d]) { ok = 0; break; }
            cz[d] = ci;
        }
        if (!ok) continue;
        size_t cell = ((size_t)cz[0] * grid[1] + cz[1]) * grid[2] + cz[2];
        int v = lookup[cell];
        if (v == -1) {
            if (num_voxels >= max_voxels) continue;   /* continue, not break (A.1) */
            v = num_voxels++;
            lookup[cell] = v;
            coords[v * 3 + 0] = cz[0];
            coords[v * 3 + 1] = cz[1];
            coords[v * 3 + 2] = cz[2];
        }
        if (num_per[v] < max_pts) {
            memcpy(voxels + ((size_t)v * max_pts + num_per[v]) * c,
                   points + (size_t)i * c, c * sizeof(float));
            num_per[v]++;
        }
    }
    for (int v = 0; v < num_voxels; ++v) {
        size_t cell = ((size_t)coords[v * 3] * grid[1] + coords[v * 3 + 1]) * grid[2] + coords[v * 3 + 2];
        lookup[cell] = -1;
    }
    return num_voxels;
}

/* ------------------------------------------------------------------------------------
 * Rotated BEV overlap.  Follows iou3d_nms_kernel.cu:42-232 (Point helpers, check_in_box2d
 * with its 1e-2 MARGIN, intersection(), angular bubble sort, shoelace), in fp32.
 * ------------------------------------------------------------------------------------ */
typedef struct { float x, y; } pt_t;
static const float ORACLE_EPS = 1e-8f;

static float cross3(pt_t p1, pt_t p2, pt_t p0) {            /* kernel.cu:46-48 */
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
static float cross2(pt_t a, pt_t b) { return a.x * b.y - a.y * b.x; }   /* :42-44 */

static int rect_cross(pt_t p1, pt_t p2, pt_t q1, pt_t q2) {  /* :50-56 */
    return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
           fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}
static int in_box2d(const float *box, pt_t p) {              /* :58-68 */
    const float MARGIN = 1e-2f;
    float cx = box[0], cy = box[1];
    float ac = cosf(-box[6]), as = sinf(-box[6]);
    float rx = (p.x - cx) * ac + (p.y - cy) * (-as);
    float ry = (p.x - cx) * as + (p.y - cy) * ac;
    return fabsf(rx) < box[3] / 2 + MARGIN && fabsf(ry) < box[4] / 2 + MARGIN;
}
static int seg_intersection(pt_t p1, pt_t p0, pt_t q1, pt_t q0, pt_t *ans) {  /* :70-100 */
    if (!rect_cross(p0, p1, q0, q1)) return 0;
    float s1 = cross3(q0, p1, p0);
    float s2 = cross3(p1, q1, p0);
    float s3 = cross3(p0, q1, q0);
    float s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    float s5 = cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > ORACLE_EPS) {
        ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        float D = a0 * b1 - a1 * b0;
        ans->x = (b0 * c1 - b1 * c0) / D;
        ans->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}
static void rot_center(pt_t c, float ac, float as, pt_t *p) {   /* :102-106 */
    float nx = (p->x - c.x) * ac + (p->y - c.y) * (-as) + c.x;
    float ny = (p->x - c.x) * as + (p->y - c.y) * ac + c.y;
    p->x = nx; p->y = ny;
}

float oracle_box_overlap(const float *a, const float *b)    /* :111-232 */
{
    float a_hx = a[3] / 2, b_hx = b[3] / 2, a_hy = a[4] / 2, b_hy = b[4] / 2;
    pt_t ca = { a[0], a[1] }, cb = { b[0], b[1] };
    pt_t A[5] = { { a[0] - a_hx, a[1] - a_hy }, { a[0] + a_hx, a[1] - a_hy },
                  { a[0] + a_hx, a[1] + a_hy }, { a[0] - a_hx, a[1] + a_hy } };
    pt_t B[5] = { { b[0] - b_hx, b[1] - b_hy }, { b[0] + b_hx, b[1] - b_hy },
                  { b[0] + b_hx, b[1] + b_hy }, { b[0] - b_hx, b[1] + b_hy } };
    float acs = cosf(a[6]), asn = sinf(a[6]), bcs = cosf(b[6]), bsn = sinf(b[6]);
    for (int k = 0; k < 4; ++k) { rot_center(ca, acs, asn, &A[k]); rot_center(cb, bcs, bsn, &B[k]); }
    A[4] = A[0]; B[4] = B[0];

    pt_t cp[16], center = { 0, 0 };
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (seg_intersection(A[i + 1], A[i], B[j + 1], B[j], &cp[cnt])) {
                center.x += cp[cnt].x; center.y += cp[cnt].y; cnt++;
            }
    for (int k = 0; k < 4; ++k) {
        if (in_box2d(a, B[k])) { center.x += B[k].x; center.y += B[k].y; cp[cnt++] = B[k]; }
        if (in_box2d(b, A[k])) { center.x += A[k].x; center.y += A[k].y; cp[cnt++] = A[k]; }
    }
    center.x /= cnt; center.y /= cnt;
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i) {
            float t0 = atan2f(cp[i].y - center.y, cp[i].x - center.x);
            float t1 = atan2f(cp[i + 1].y - center.y, cp[i + 1].x - center.x);
            if (t0 > t1) { pt_t t = cp[i]; cp[i] = cp[i + 1]; cp[i + 1] = t; }
        }
    float area = 0;
    for (int k = 0; k < cnt - 1; ++k) {
        pt_t u = { cp[k].x - cp[0].x, cp[k].y - cp[0].y };
        pt_t v = { cp[k + 1].x - cp[0].x, cp[k + 1].y - cp[0].y };
        area += cross2(u, v);
    }
    return fabsf(area) / 2.0f;
}

float oracle_iou_bev(const float *a, const float *b)         /* :328-335 */
{
    float sa = a[3] * a[4], sb = b[3] * b[4];
    float so = oracle_box_overlap(a, b);
    return so / fmaxf(sa + sb - so, ORACLE_EPS);
}

void oracle_boxes_iou_bev(const float *boxes_a, int na, const float *boxes_b, int nb, float *out)
{
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j)
            out[(size_t)i * nb + j] = oracle_iou_bev(boxes_a + i * 7, boxes_b + j * 7);
}

/* boxes_overlap_bev_gpu (iou3d_nms_kernel.cu:337-350): the overlap AREA matrix the tracker's overlap filter and
 * boxes_iou3d_gpu (iou3d_nms_utils.py:74-107) are built on */
void oracle_boxes_overlap_bev(const float *boxes_a, int na, const float *boxes_b, int nb, float *out)
{
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j)
            out[(size_t)i * nb + j] = oracle_box_overlap(boxes_a + i * 7, boxes_b + j * 7);
}

/* points_in_boxes_gpu_v2 (roiaware_pool3d_kernel.cu:16-36,352-372): mask[t][m] = 1 iff point m is inside box t;
 * GPU margin 1e-5 on x/y in the box frame, none on z */
void oracle_points_in_boxes(const float *pts, int n_pts, int pt_stride, const float *boxes, int n_boxes, int32_t *mask)
{
    for (int t = 0; t < n_boxes; ++t) {
        const float *b = boxes + (size_t)t * 7;
        float cosa = cosf(-b[6]), sina = sinf(-b[6]);
        for (int m = 0; m < n_pts; ++m) {
            const float *q = pts + (size_t)m * pt_stride;
            int in = 0;
            if (!(fabsf(q[2] - b[2]) > b[5] / 2.0f)) {
                float sx = q[0] - b[0], sy = q[1] - b[1];
                float lx = sx * cosa + sy * (-sina), ly = sx * sina + sy * cosa;
                in = (fabsf(lx) < b[3] / 2.0f + 1e-5f) & (fabsf(ly) < b[4] / 2.0f + 1e-5f);
            }
            mask[(size_t)t * n_pts + m] = in;
        }
    }
}

/* NMS over boxes ALREADY sorted by descending score (iou3d_nms_utils.py:154-170 sorts, then
 * nms_kernel builds the j>i suppression mask (:386-430) and the host scans it serially
 * (iou3d_nms.cpp:139-157)).  keep[] receives indices into the sorted order; returns count. */
int oracle_nms_bev(const float *boxes, int n, float thresh, int64_t *keep)
{
    unsigned char *removed = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (removed[i]) continue;
        keep[nk++] = i;
        for (int j = i + 1; j < n; ++j)
            if (!removed[j] && oracle_iou_bev(boxes + i * 7, boxes + j * 7) > thresh) removed[j] = 1;
    }
    free(removed);
    return nk;
}

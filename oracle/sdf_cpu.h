/*
 * oracle/sdf_cpu.h — CPU evaluation of hpsdf_sdf_program (include/hpsdf.h).
 *
 * TEST INFRASTRUCTURE ONLY. This is the function F handed to the reference's Octree::Create
 * (Include/HP/Octree.h:50) by oracle/ref_driver.cpp and to the C restatement oracle/hp_oracle.c;
 * the product evaluates the same programs on the device (csrc/sdf_eval.cuh) and never includes this.
 *
 * Closed forms are the usual f64 ones (sphere, box, torus, capsule after Quilez); 3-term sums
 * associate as a0 + (a1 + a2), the order Eigen's fixed-size reductions use.
 */
#ifndef HPSDF_ORACLE_SDF_CPU_H
#define HPSDF_ORACLE_SDF_CPU_H

#include <math.h>
#include <stdint.h>
#include "../include/hpsdf.h"

/* MESH / OCTREE primitives are resolved by the embedding driver. */
typedef double (*hporacle_ext_eval)(uint32_t op, const void* handle, const double x[3]);

static inline double hporacle_len3(double x, double y, double z) { return sqrt(x * x + (y * y + z * z)); }

static inline double hporacle_prim(const hpsdf_sdf_instr* in, const double x[3], hporacle_ext_eval ext)
{
    const double* p = in->p;
    switch (in->op)
    {
        case HPSDF_PRIM_SPHERE:
            return hporacle_len3(x[0] - p[0], x[1] - p[1], x[2] - p[2]) - p[3];
        case HPSDF_PRIM_BOX:
        {
            const double qx = fabs(x[0] - p[0]) - p[3], qy = fabs(x[1] - p[1]) - p[4], qz = fabs(x[2] - p[2]) - p[5];
            const double mx = fmax(qx, 0.0), my = fmax(qy, 0.0), mz = fmax(qz, 0.0);
            return hporacle_len3(mx, my, mz) + fmin(fmax(qx, fmax(qy, qz)), 0.0);
        }
        case HPSDF_PRIM_TORUS:
        {
            const int a = (int)p[5];
            const double d[3] = { x[0] - p[0], x[1] - p[1], x[2] - p[2] };
            const double u = d[(a + 1) % 3], v = d[(a + 2) % 3], h = d[a];
            const double q = sqrt(u * u + v * v) - p[3];
            return sqrt(q * q + h * h) - p[4];
        }
        case HPSDF_PRIM_CAPSULE:
        {
            const double pax = x[0] - p[0], pay = x[1] - p[1], paz = x[2] - p[2];
            const double bax = p[3] - p[0], bay = p[4] - p[1], baz = p[5] - p[2];
            double h = (pax * bax + (pay * bay + paz * baz)) / (bax * bax + (bay * bay + baz * baz));
            h = fmin(fmax(h, 0.0), 1.0);
            return hporacle_len3(pax - bax * h, pay - bay * h, paz - baz * h) - p[6];
        }
        case HPSDF_PRIM_PLANE:
            return (p[0] * x[0] + (p[1] * x[1] + p[2] * x[2])) - p[3];
        case HPSDF_PRIM_MESH:
        case HPSDF_PRIM_OCTREE:
            return ext ? ext(in->op, in->handle, x) : NAN;
        default:
            return NAN;
    }
}

static inline double hporacle_sdf_eval(const hpsdf_sdf_instr* instr, uint32_t n, const double x[3], hporacle_ext_eval ext)
{
    double st[HPSDF_PROGRAM_MAX_STACK];
    int sp = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        const uint32_t op = instr[i].op;
        if (op < HPSDF_OP_UNION) { st[sp++] = hporacle_prim(&instr[i], x, ext); continue; }
        if (op == HPSDF_OP_NEGATE) { st[sp - 1] = -st[sp - 1]; continue; }
        const double b = st[--sp], a = st[sp - 1];
        st[sp - 1] = op == HPSDF_OP_UNION ? fmin(a, b) : op == HPSDF_OP_INTERSECT ? fmax(a, b) : fmax(a, -b);
    }
    return st[0];
}

#endif

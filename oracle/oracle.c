/*
 * oracle.c — CPU restatement of AffAction's action-feasibility predictor.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED against the real
 * Rcs/Tropic-linked predictor; the CUDA path is checked
 * against THIS file.
 *
 * What follows the reference verbatim (all in /root/reference/src):
 *   predict loop ............ TrajectoryPredictor.cpp:79-406
 *   computeIK ............... TrajectoryPredictor.cpp:573-753
 *   elbow / wrist null space  TrajectoryPredictor.cpp:448-570
 *   checkState .............. TrajectoryPredictor.cpp:756-886
 *   collision toggling ...... CollisionModelConstraint.h:101-147
 * What restates un-vendored Rcs / Tropic behaviour:
 *   forward kinematics, task values / Jacobians, joint-limit cost, the damped
 *   weighted right inverse, via-point polynomials, activation / blending,
 *   ConnectBodyConstraint, broad phase and shape distances.  DESIGN.md lists
 *   each rule that had to be chosen.
 *
 * Arithmetic discipline: every multi-term expression is written in the exact
 * order (and with the explicit fma()s) the CUDA kernel uses, transcendental
 * functions come from include/aff_detmath.h, and this file is compiled with
 * -ffp-contract=off, so results are comparable bit for bit.
 * Build with -DAFF_ORACLE_LIBM to use glibc sin/cos/atan2/acos instead (used
 * to show the verdicts do not depend on that substitution).
 */
#include "oracle.h"
#include "aff_detmath.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef AFF_ORACLE_LIBM
#define o_sincos(x, s, c) do { *(s) = sin(x); *(c) = cos(x); } while (0)
#define o_sin(x) sin(x)
#define o_atan2(y, x) atan2(y, x)
static double o_acos(double x) { return x >= 1.0 ? 0.0 : (x <= -1.0 ? M_PI : acos(x)); }
int oracle_uses_libm(void) { return 1; }
#else
#define o_sincos(x, s, c) aff_sincos(x, s, c)
#define o_sin(x) aff_sin(x)
#define o_atan2(y, x) aff_atan2(y, x)
#define o_acos(x) aff_acos(x)
int oracle_uses_libm(void) { return 0; }
#endif

#define MAX_TRACKING_ERROR 0.05     /* TrajectoryPredictor.cpp:50  */
#define TRAJ_EPS 1.0e-8             /* TRAJECTORY1D_ALMOST_ZERO    */
#define TRAJ_HORIZON 1.0            /* ActionBase.cpp:137          */
#define MAX_VIA_UNKNOWNS 12
#define NO_LIMIT 1.0e300

typedef struct { double org[3]; double rot[3][3]; } HTr;

/* AFF_ORACLE_TRACE=n prints the task state every n steps (debugging aid) */
static int g_trace = -1;

/* ------------------------------------------------------------------ HTr ops */
static void htr_identity(HTr* A)
{
  memset(A, 0, sizeof(*A));
  A->rot[0][0] = A->rot[1][1] = A->rot[2][2] = 1.0;
}

static void htr_from12(HTr* A, const double* v)
{
  for (int i = 0; i < 3; ++i) A->org[i] = v[i];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A->rot[i][j] = v[3 + 3 * i + j];
}

static void htr_to12(const HTr* A, double* v)
{
  for (int i = 0; i < 3; ++i) v[i] = A->org[i];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) v[3 + 3 * i + j] = A->rot[i][j];
}

/* out = T o P : frame T given relative to P (HTr_transform). */
static void htr_compose(HTr* out, const HTr* T, const HTr* P)
{
  HTr r;
  for (int i = 0; i < 3; ++i)
    r.org[i] = fma(P->rot[2][i], T->org[2], fma(P->rot[1][i], T->org[1], fma(P->rot[0][i], T->org[0], P->org[i])));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      r.rot[i][j] = fma(T->rot[i][2], P->rot[2][j], fma(T->rot[i][1], P->rot[1][j], T->rot[i][0] * P->rot[0][j]));
  *out = r;
}

/* Elementary joint motion applied to the joint frame. */
static void htr_apply_joint(HTr* A, int type, double q)
{
  if (type < AFF_JNT_ROT_X) {
    const int d = type;
    for (int k = 0; k < 3; ++k) A->org[k] = fma(q, A->rot[d][k], A->org[k]);
  } else {
    const int d = type - AFF_JNT_ROT_X, a = (d + 1) % 3, b = (d + 2) % 3;
    double s, c;
    o_sincos(q, &s, &c);
    for (int k = 0; k < 3; ++k) {
      const double ra = A->rot[a][k], rb = A->rot[b][k];
      A->rot[a][k] = fma(s, rb, c * ra);
      A->rot[b][k] = fma(c, rb, -(s * ra));
    }
  }
}

/* A_rel such that child = A_rel o parent  (HTr_invTransform). */
static void htr_relative(HTr* rel, const HTr* child, const HTr* parent)
{
  double d[3];
  for (int k = 0; k < 3; ++k) d[k] = child->org[k] - parent->org[k];
  for (int i = 0; i < 3; ++i)
    rel->org[i] = fma(parent->rot[i][2], d[2], fma(parent->rot[i][1], d[1], parent->rot[i][0] * d[0]));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      rel->rot[i][j] = fma(child->rot[i][2], parent->rot[j][2],
                           fma(child->rot[i][1], parent->rot[j][1], child->rot[i][0] * parent->rot[j][0]));
}

static void v3_cross(double* o, const double* a, const double* b)
{
  o[0] = fma(a[1], b[2], -(a[2] * b[1]));
  o[1] = fma(a[2], b[0], -(a[0] * b[2]));
  o[2] = fma(a[0], b[1], -(a[1] * b[0]));
}

static double v3_dot(const double* a, const double* b)
{
  return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0]));
}

/* o = R * v */
static void m3_mulv(double* o, const double R[3][3], const double* v)
{
  for (int i = 0; i < 3; ++i) o[i] = fma(R[i][2], v[2], fma(R[i][1], v[1], R[i][0] * v[0]));
}

static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------- graph */
typedef struct {
  const aff_scene_desc* d;
  int nB, nQ, nJ, nS;
  int* jacobi;       /* [nQ]  jacobi index or -1                    */
  int* ikj;          /* [nJ]  joint id                              */
  int* jbody;        /* [nQ]  owning body                           */
  int* parent;       /* [nB]  mutable (ConnectBodyConstraint)        */
  unsigned char* has_rel; HTr* rel; /* pose of a re-parented body    */
  unsigned char* sdist;   /* [nS] mutable RCSSHAPE_COMPUTE_DISTANCE  */
  double* q;         /* [nQ]                                        */
  double* qdot;      /* [nQ]                                        */
  HTr* A_BI;         /* [nB]                                        */
  HTr* A_JI;         /* [nQ]                                        */
  int* order;        /* [nB] parents before children                */
  int* sbody; int nSB;    /* bodies owning >= 1 primitive shape       */
  int* tree_of;      /* [nB] scratch                                */
  double* aabb;      /* [nB*6] scratch                              */
  unsigned char* explicit_bp; /* [nB]                                */
} Graph;

static void graph_free(Graph* g)
{
  free(g->jacobi); free(g->ikj); free(g->jbody); free(g->parent); free(g->has_rel); free(g->rel);
  free(g->sdist); free(g->q); free(g->qdot); free(g->A_BI); free(g->A_JI); free(g->order);
  free(g->sbody); free(g->tree_of); free(g->aabb); free(g->explicit_bp);
}

static void graph_sort(Graph* g)
{
  /* topological order: repeatedly emit bodies whose parent is emitted; stable in id */
  unsigned char* done = (unsigned char*)calloc(g->nB, 1);
  int n = 0;
  while (n < g->nB) {
    int progressed = 0;
    for (int b = 0; b < g->nB; ++b) {
      if (done[b]) continue;
      const int p = g->parent[b];
      if (p < 0 || done[p]) { done[b] = 1; g->order[n++] = b; progressed = 1; }
    }
    if (!progressed) break; /* cycle: malformed */
  }
  free(done);
}

static int graph_init(Graph* g, const aff_scene_desc* d, const aff_candidate* cand)
{
  memset(g, 0, sizeof(*g));
  g->d = d; g->nB = d->n_bodies; g->nQ = d->n_joints; g->nS = d->n_shapes;
  g->jacobi = (int*)malloc(sizeof(int) * (g->nQ + 1));
  g->jbody = (int*)malloc(sizeof(int) * (g->nQ + 1));
  g->ikj = (int*)malloc(sizeof(int) * (g->nQ + 1));
  g->parent = (int*)malloc(sizeof(int) * g->nB);
  g->has_rel = (unsigned char*)calloc(g->nB, 1);
  g->rel = (HTr*)calloc(g->nB, sizeof(HTr));
  g->sdist = (unsigned char*)malloc(g->nS + 1);
  g->q = (double*)calloc(g->nQ + 1, sizeof(double));
  g->qdot = (double*)calloc(g->nQ + 1, sizeof(double));
  g->A_BI = (HTr*)calloc(g->nB, sizeof(HTr));
  g->A_JI = (HTr*)calloc(g->nQ + 1, sizeof(HTr));
  g->order = (int*)malloc(sizeof(int) * g->nB);
  g->sbody = (int*)malloc(sizeof(int) * g->nB);
  g->tree_of = (int*)malloc(sizeof(int) * g->nB);
  g->aabb = (double*)malloc(sizeof(double) * 6 * g->nB);
  g->explicit_bp = (unsigned char*)calloc(g->nB, 1);
  g->nJ = 0;
  for (int j = 0; j < g->nQ; ++j) {
    if (d->joint_constrained[j]) g->jacobi[j] = -1;
    else { g->jacobi[j] = g->nJ; g->ikj[g->nJ++] = j; }
    g->q[j] = d->joint_q_init[j];
  }
  for (int b = 0; b < g->nB; ++b) {
    g->parent[b] = d->body_parent[b];
    for (int k = 0; k < d->body_n_joints[b]; ++k) g->jbody[d->body_first_joint[b] + k] = b;
    if (d->body_n_shapes[b] > 0) g->sbody[g->nSB++] = b;
  }
  for (int s = 0; s < g->nS; ++s) g->sdist[s] = d->shape_distance[s];
  for (int i = 0; i < d->n_bp_bodies; ++i) g->explicit_bp[d->bp_body[i]] = 1;
  if (cand && cand->pose_body >= 0) {
    const int b = cand->pose_body;
    if (b >= g->nB || !d->body_rbj[b] || d->body_n_joints[b] < 6) return -1;
    for (int k = 0; k < 6; ++k) g->q[d->body_first_joint[b] + k] = cand->pose_q6[k];
  }
  graph_sort(g);
  return 0;
}

/* RcsGraph_setState: forward kinematics over every body. */
static void graph_fk(Graph* g)
{
  const aff_scene_desc* d = g->d;
  HTr I; htr_identity(&I);
  for (int n = 0; n < g->nB; ++n) {
    const int b = g->order[n];
    HTr cur = (g->parent[b] >= 0) ? g->A_BI[g->parent[b]] : I;
    const int j0 = d->body_first_joint[b], nj = d->body_n_joints[b];
    if (g->has_rel[b]) {
      htr_compose(&cur, &g->rel[b], &cur);
      for (int k = 0; k < nj; ++k) g->A_JI[j0 + k] = cur;
    } else {
      for (int k = 0; k < nj; ++k) {
        HTr T; htr_from12(&T, d->joint_A_JP + 12 * (j0 + k));
        htr_compose(&cur, &T, &cur);
        htr_apply_joint(&cur, d->joint_type[j0 + k], g->q[j0 + k]);
        g->A_JI[j0 + k] = cur;
      }
    }
    HTr T; htr_from12(&T, d->body_A_BP + 12 * b);
    htr_compose(&g->A_BI[b], &T, &cur);
  }
}

/* Is IK joint `jid` on the chain from `body` to the root? */
static int joint_moves_body(const Graph* g, int jid, int body)
{
  const int jb = g->jbody[jid];
  for (int b = body; b >= 0; b = g->parent[b]) if (b == jb) return !g->has_rel[b];
  return 0;
}

/* tropic::ConnectBodyConstraint: re-parent `child` under `parent` keeping its
 * world pose; the six rigid-body dofs are replaced by the relative transform. */
static int graph_connect(Graph* g, int child, int parent, int identity)
{
  if (child < 0 || parent < 0 || child >= g->nB || parent >= g->nB || !g->d->body_rbj[child]) return -1;
  /* frame at the end of the rigid-body joints = body frame when A_BP is identity */
  HTr rel;
  htr_relative(&rel, &g->A_BI[child], &g->A_BI[parent]);
  if (identity) {   /* cbc->setConnectTransform(HTr_identity()), e.g. src/ActionDrop.cpp:237-239 */
    for (int i = 0; i < 3; ++i) { rel.org[i] = 0.0; for (int j = 0; j < 3; ++j) rel.rot[i][j] = (i == j) ? 1.0 : 0.0; }
  }
  g->rel[child] = rel;
  g->has_rel[child] = 1;
  g->parent[child] = parent;
  graph_sort(g);
  return 0;
}

/* ---------------------------------------------------------------- Jacobians */
/* Translation / rotation Jacobian columns of IK joint jid for a point p fixed
 * to `body`, in world coordinates. */
static void jac_cols(const Graph* g, int jid, int body, const double* p, double* colT, double* colR)
{
  colT[0] = colT[1] = colT[2] = 0.0;
  colR[0] = colR[1] = colR[2] = 0.0;
  if (body < 0 || !joint_moves_body(g, jid, body)) return;
  const int type = g->d->joint_type[jid];
  const HTr* A = &g->A_JI[jid];
  if (type < AFF_JNT_ROT_X) {
    for (int k = 0; k < 3; ++k) colT[k] = A->rot[type][k];
  } else {
    const double* ax = A->rot[type - AFF_JNT_ROT_X];
    double r[3];
    for (int k = 0; k < 3; ++k) r[k] = p[k] - A->org[k];
    v3_cross(colT, ax, r);
    for (int k = 0; k < 3; ++k) colR[k] = ax[k];
  }
}

/* ------------------------------------------------------------------- tasks */
static int task_dim(const aff_task_desc* t)
{
  switch (t->kind) {
    case AFF_TASK_XYZ: return 3;
    case AFF_TASK_X: case AFF_TASK_Y: case AFF_TASK_Z: case AFF_TASK_INCLINATION: return 1;
    case AFF_TASK_POLAR: return 2;
    case AFF_TASK_JOINTS: return t->n_joints;
    default: return -1;
  }
}

static void polar_axis(const Graph* g, const aff_task_desc* t, double* a)
{
  const double* aI = g->A_BI[t->effector].rot[t->axis];
  if (t->ref_body >= 0) m3_mulv(a, g->A_BI[t->ref_body].rot, aI);
  else { a[0] = aI[0]; a[1] = aI[1]; a[2] = aI[2]; }
}

/* Orthonormal tangent basis (e1, e2) with e1 x e2 = a. */
static void tangent_basis(const double* a, double* e1, double* e2)
{
  int k = 0;
  double m = fabs(a[0]);
  if (fabs(a[1]) < m) { k = 1; m = fabs(a[1]); }
  if (fabs(a[2]) < m) { k = 2; }
  double t[3];
  for (int i = 0; i < 3; ++i) t[i] = -(a[k] * a[i]);
  t[k] = t[k] + 1.0;
  const double n = sqrt(v3_dot(t, t));
  for (int i = 0; i < 3; ++i) e1[i] = t[i] / n;
  v3_cross(e2, a, e1);
}

static void inclination_axes(const Graph* g, const aff_task_desc* t, double* u, double* v)
{
  const double* aI = g->A_BI[t->effector].rot[t->axis];
  u[0] = aI[0]; u[1] = aI[1]; u[2] = aI[2];
  if (t->ref_body >= 0) { const double* z = g->A_BI[t->ref_body].rot[2]; v[0] = z[0]; v[1] = z[1]; v[2] = z[2]; }
  else { v[0] = 0.0; v[1] = 0.0; v[2] = 1.0; }
}

/* Task::computeX */
static void task_x(const Graph* g, const aff_task_desc* t, double* x)
{
  switch (t->kind) {
    case AFF_TASK_XYZ: case AFF_TASK_X: case AFF_TASK_Y: case AFF_TASK_Z: {
      double r[3], xr[3];
      for (int k = 0; k < 3; ++k)
        r[k] = (t->ref_body >= 0) ? g->A_BI[t->effector].org[k] - g->A_BI[t->ref_body].org[k] : g->A_BI[t->effector].org[k];
      const int fr = (t->ref_frame == -2) ? -1 : ((t->ref_frame >= 0) ? t->ref_frame : t->ref_body);
      if (fr >= 0) m3_mulv(xr, g->A_BI[fr].rot, r);
      else { xr[0] = r[0]; xr[1] = r[1]; xr[2] = r[2]; }
      if (t->kind == AFF_TASK_XYZ) { x[0] = xr[0]; x[1] = xr[1]; x[2] = xr[2]; }
      else x[0] = xr[t->kind - AFF_TASK_X];
      break;
    }
    case AFF_TASK_POLAR: {
      double a[3];
      polar_axis(g, t, a);
      x[0] = o_acos(a[2]);
      x[1] = o_atan2(a[1], a[0]);
      break;
    }
    case AFF_TASK_INCLINATION: {
      double u[3], v[3];
      inclination_axes(g, t, u, v);
      x[0] = o_acos(v3_dot(u, v));
      break;
    }
    case AFF_TASK_JOINTS:
      for (int i = 0; i < t->n_joints; ++i) x[i] = g->q[t->joints[i]];
      break;
    default: break;
  }
}

static double region_dx(const aff_task_desc* t, double e)
{
  if (!t->has_region) return e;
  const double off = -e;   /* x_curr - x_des */
  if (off > t->region_max) return e + t->region_max;
  if (off < t->region_min) return e + t->region_min;
  return t->region_dx_scaling * e;
}

/* Task::computeDX : task-space error towards x_des at the current state. */
static void task_dx(const Graph* g, const aff_task_desc* t, const double* x_des, double* dx)
{
  if (t->kind == AFF_TASK_POLAR) {
    double a[3], ad[3], e1[3], e2[3], n[3], w[3];
    polar_axis(g, t, a);
    tangent_basis(a, e1, e2);
    double sp, cp, st, ct;
    o_sincos(x_des[0], &sp, &cp);
    o_sincos(x_des[1], &st, &ct);
    ad[0] = sp * ct; ad[1] = sp * st; ad[2] = cp;
    v3_cross(n, a, ad);
    const double s = sqrt(v3_dot(n, n));
    const double c = v3_dot(a, ad);
    const double ang = o_atan2(s, c);
    if (s > 1.0e-12) { const double f = ang / s; for (int k = 0; k < 3; ++k) w[k] = n[k] * f; }
    else { w[0] = w[1] = w[2] = 0.0; }
    dx[0] = v3_dot(e1, w);
    dx[1] = v3_dot(e2, w);
    return;
  }
  const int dim = task_dim(t);
  double x[AFF_MAX_TASK_JOINTS];
  task_x(g, t, x);
  for (int i = 0; i < dim; ++i) dx[i] = region_dx(t, x_des[i] - x[i]);
}

/* Task::computeJ : rows appended to J (row-major, nJ columns). */
static void task_J(const Graph* g, const aff_task_desc* t, double* J)
{
  const int nJ = g->nJ, dim = task_dim(t);
  memset(J, 0, sizeof(double) * (size_t)dim * nJ);
  switch (t->kind) {
    case AFF_TASK_XYZ: case AFF_TASK_X: case AFF_TASK_Y: case AFF_TASK_Z: {
      const int fr = (t->ref_frame == -2) ? -1 : ((t->ref_frame >= 0) ? t->ref_frame : t->ref_body);
      const double* pe = g->A_BI[t->effector].org;
      double zero[3] = {0.0, 0.0, 0.0};
      const double* pr = (t->ref_body >= 0) ? g->A_BI[t->ref_body].org : zero;
      double r12[3];
      for (int k = 0; k < 3; ++k) r12[k] = (t->ref_body >= 0) ? pe[k] - pr[k] : pe[k];
      for (int c = 0; c < nJ; ++c) {
        const int jid = g->ikj[c];
        double te[3], re[3], tr[3], rr[3], tf[3], rf[3], cr[3], col[3], out[3];
        jac_cols(g, jid, t->effector, pe, te, re);
        jac_cols(g, jid, t->ref_body, pr, tr, rr);
        jac_cols(g, jid, fr, fr >= 0 ? g->A_BI[fr].org : zero, tf, rf);
        v3_cross(cr, r12, rf);
        for (int k = 0; k < 3; ++k) col[k] = (te[k] - tr[k]) + cr[k];
        if (fr >= 0) m3_mulv(out, g->A_BI[fr].rot, col);
        else { out[0] = col[0]; out[1] = col[1]; out[2] = col[2]; }
        if (t->kind == AFF_TASK_XYZ) { J[c] = out[0]; J[nJ + c] = out[1]; J[2 * nJ + c] = out[2]; }
        else J[c] = out[t->kind - AFF_TASK_X];
      }
      break;
    }
    case AFF_TASK_POLAR: {
      double a[3], e1[3], e2[3];
      polar_axis(g, t, a);
      tangent_basis(a, e1, e2);
      const double* pe = g->A_BI[t->effector].org;
      for (int c = 0; c < nJ; ++c) {
        const int jid = g->ikj[c];
        double te[3], re[3], tr[3], rr[3], wI[3], w[3];
        jac_cols(g, jid, t->effector, pe, te, re);
        jac_cols(g, jid, t->ref_body, pe, tr, rr);
        for (int k = 0; k < 3; ++k) wI[k] = re[k] - rr[k];
        if (t->ref_body >= 0) m3_mulv(w, g->A_BI[t->ref_body].rot, wI);
        else { w[0] = wI[0]; w[1] = wI[1]; w[2] = wI[2]; }
        J[c] = v3_dot(e1, w);
        J[nJ + c] = v3_dot(e2, w);
      }
      break;
    }
    case AFF_TASK_INCLINATION: {
      double u[3], v[3], n[3];
      inclination_axes(g, t, u, v);
      v3_cross(n, v, u);
      const double sn = sqrt(v3_dot(n, n));
      if (sn > 1.0e-8) {
        for (int k = 0; k < 3; ++k) n[k] = n[k] / sn;
        const double* pe = g->A_BI[t->effector].org;
        for (int c = 0; c < nJ; ++c) {
          const int jid = g->ikj[c];
          double te[3], re[3], tr[3], rr[3], wI[3];
          jac_cols(g, jid, t->effector, pe, te, re);
          jac_cols(g, jid, t->ref_body, pe, tr, rr);
          for (int k = 0; k < 3; ++k) wI[k] = re[k] - rr[k];
          J[c] = v3_dot(n, wI);
        }
      }
      break;
    }
    case AFF_TASK_JOINTS:
      for (int i = 0; i < t->n_joints; ++i) {
        const int c = g->jacobi[t->joints[i]];
        if (c >= 0) J[i * nJ + c] = 1.0;
      }
      break;
    default: break;
  }
}

/* ------------------------------------------------- via-point trajectories */
/* Solve the K x K system in place (Gaussian elimination, partial pivoting,
 * first maximal pivot).  A is row-major with stride K+1 (last column = rhs). */
static void gauss_solve(double* A, int K, double* sol)
{
  const int S = K + 1;
  for (int c = 0; c < K; ++c) {
    int piv = c;
    double best = fabs(A[c * S + c]);
    for (int r = c + 1; r < K; ++r) { const double v = fabs(A[r * S + c]); if (v > best) { best = v; piv = r; } }
    if (piv != c) for (int k = c; k < S; ++k) { const double tmp = A[c * S + k]; A[c * S + k] = A[piv * S + k]; A[piv * S + k] = tmp; }
    const double p = A[c * S + c];
    if (p == 0.0) continue; /* singular: leave the column */
    for (int r = c + 1; r < K; ++r) {
      const double f = A[r * S + c] / p;
      for (int k = c + 1; k < S; ++k) A[r * S + k] = fma(-f, A[c * S + k], A[r * S + k]);
    }
  }
  for (int r = K - 1; r >= 0; --r) {
    double s = A[r * S + K];
    for (int k = r + 1; k < K; ++k) s = fma(-A[r * S + k], sol[k], s);
    const double p = A[r * S + r];
    sol[r] = (p != 0.0) ? s / p : 0.0;
  }
}

/* One receding-horizon step of a ViaPointTrajectory1D.
 * vias: rows (t_rel, x, xd, xdd, flag) sorted by time, all t_rel > TRAJ_EPS.
 * The window runs up to and including the first full (flag 7) constraint at or
 * beyond the horizon (else the last full one, else the last constraint); the
 * minimal-degree polynomial through the current state and the window is
 * evaluated at dt.  Time is normalised by the window length. */
void oracle_via_step(double st[3], const double* vias, int n, double horizon, double dt)
{
  if (n <= 0) { st[1] = 0.0; st[2] = 0.0; return; }
  int G = -1, lastFull = -1;
  for (int i = 0; i < n; ++i) {
    const int flag = (int)vias[5 * i + 4];
    if (flag == 7) { lastFull = i; if (vias[5 * i] >= horizon) { G = i; break; } }
  }
  if (G < 0) G = (lastFull >= 0) ? lastFull : n - 1;
  int K = 0, nwin = 0;
  for (int i = 0; i <= G; ++i) {
    const int flag = (int)vias[5 * i + 4];
    const int add = (flag & 1) + ((flag >> 1) & 1) + ((flag >> 2) & 1);
    if (K + add > MAX_VIA_UNKNOWNS) break;
    K += add; nwin = i + 1;
  }
  if (K == 0) { st[1] = 0.0; st[2] = 0.0; return; }
  const double T = vias[5 * (nwin - 1)];
  const double c0 = st[0], c1 = st[1] * T, c2 = (0.5 * st[2]) * (T * T);
  double A[MAX_VIA_UNKNOWNS * (MAX_VIA_UNKNOWNS + 1)], c[MAX_VIA_UNKNOWNS];
  const int S = K + 1;
  int row = 0;
  for (int i = 0; i < nwin; ++i) {
    const double* v = vias + 5 * i;
    const int flag = (int)v[4];
    const double tau = v[0] / T;
    /* powers tau^0 .. tau^(K+2) */
    double pw[MAX_VIA_UNKNOWNS + 3];
    pw[0] = 1.0;
    for (int k = 1; k < K + 3; ++k) pw[k] = pw[k - 1] * tau;
    if (flag & 1) {
      for (int k = 0; k < K; ++k) A[row * S + k] = pw[k + 3];
      A[row * S + K] = v[1] - fma(c2, pw[2], fma(c1, tau, c0));
      ++row;
    }
    if (flag & 2) {
      for (int k = 0; k < K; ++k) A[row * S + k] = (double)(k + 3) * pw[k + 2];
      A[row * S + K] = v[2] * T - fma(2.0 * c2, tau, c1);
      ++row;
    }
    if (flag & 4) {
      for (int k = 0; k < K; ++k) A[row * S + k] = (double)((k + 3) * (k + 2)) * pw[k + 1];
      A[row * S + K] = v[3] * (T * T) - 2.0 * c2;
      ++row;
    }
  }
  gauss_solve(A, K, c);
  /* a constraint closer than dt is reached in this step: evaluate at it */
  const double s = (dt < T) ? dt / T : 1.0;
  /* Horner from the highest coefficient */
  double p = 0.0, pd = 0.0, pdd = 0.0;
  for (int k = K - 1; k >= 0; --k) {
    const int deg = k + 3;
    p = fma(p, s, c[k]);
    pd = fma(pd, s, (double)deg * c[k]);
    pdd = fma(pdd, s, (double)(deg * (deg - 1)) * c[k]);
  }
  /* p, pd, pdd are the series divided by s^3, s^2, s^1 resp. */
  const double s2 = s * s, s3 = s2 * s;
  st[0] = fma(p, s3, fma(c2, s2, fma(c1, s, c0)));
  st[1] = fma(pd, s2, fma(2.0 * c2, s, c1)) / T;
  st[2] = fma(pdd, s, 2.0 * c2) / (T * T);
}

/* -------------------------------------------------------------- IK solver */
/* Rcs::IkSolverRMR::solveRightInverse, evaluated as
 *   v = W^-1 dH,  y = (J W^-1 J^T + lambda I)^-1 (dx + J v),  dq = W^-1 J^T y - v
 * which equals J# dx - (I - J# J) W^-1 dH. */
int oracle_right_inverse(int m, int n, const double* J, const double* w,
                         const double* dx, const double* dH, double lambda, double* dq)
{
  double* v = (double*)malloc(sizeof(double) * (size_t)(n + 1));
  double* L = (double*)calloc((size_t)(m * m + 1), sizeof(double));
  double* y = (double*)malloc(sizeof(double) * (size_t)(m + 1));
  int ok = 1;
  for (int j = 0; j < n; ++j) v[j] = w[j] * dH[j];
  for (int i = 0; i < m && ok; ++i) {
    double acc = dx[i];
    for (int j = 0; j < n; ++j) acc = fma(J[i * n + j], v[j], acc);
    y[i] = acc;
  }
  for (int i = 0; i < m; ++i)
    for (int k = 0; k <= i; ++k) {
      double acc = 0.0;
      for (int j = 0; j < n; ++j) acc = fma(J[i * n + j] * w[j], J[k * n + j], acc);
      if (i == k) acc = acc + lambda;
      L[i * m + k] = acc;
    }
  /* Cholesky with reciprocal pivots: L[i][j] = t * (1 / L[j][j]); the substitutions multiply by the
   * stored reciprocals as well (one division per column instead of one per entry). */
  double* invd = (double*)malloc(sizeof(double) * (size_t)(m + 1));
  for (int j = 0; j < m && ok; ++j) {
    double s = L[j * m + j];
    for (int k = 0; k < j; ++k) s = fma(-L[j * m + k], L[j * m + k], s);
    if (!(s > 0.0)) { ok = 0; break; }
    const double ljj = sqrt(s);
    const double inv = 1.0 / ljj;
    L[j * m + j] = ljj;
    invd[j] = inv;
    for (int i = j + 1; i < m; ++i) {
      double t = L[i * m + j];
      for (int k = 0; k < j; ++k) t = fma(-L[i * m + k], L[j * m + k], t);
      L[i * m + j] = t * inv;
    }
  }
  if (ok) {
    for (int i = 0; i < m; ++i) {
      double s = y[i];
      for (int k = 0; k < i; ++k) s = fma(-L[i * m + k], y[k], s);
      y[i] = s * invd[i];
    }
    for (int i = m - 1; i >= 0; --i) {
      double s = y[i];
      for (int k = m - 1; k > i; --k) s = fma(-L[k * m + i], y[k], s);   /* column-oriented order */
      y[i] = s * invd[i];
    }
    for (int j = 0; j < n; ++j) {
      double acc = 0.0;
      for (int i = 0; i < m; ++i) acc = fma(J[i * n + j], y[i], acc);
      dq[j] = fma(w[j], acc, -v[j]);
    }
  }
  free(invd);
  free(v); free(L); free(y);
  return ok;
}

/* -------------------------------------------------------- shape distances */
static double clamp01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }

/* distance between segments p1 + s d1 and p2 + t d2 (s,t in [0,1]) */
static double seg_seg_dist(const double* p1, const double* d1, const double* p2, const double* d2)
{
  const double EPS = 1.0e-18;
  double r[3];
  for (int k = 0; k < 3; ++k) r[k] = p1[k] - p2[k];
  const double a = v3_dot(d1, d1), e = v3_dot(d2, d2), f = v3_dot(d2, r);
  double s, t;
  if (a <= EPS && e <= EPS) { s = 0.0; t = 0.0; }
  else if (a <= EPS) { s = 0.0; t = clamp01(f / e); }
  else {
    const double c = v3_dot(d1, r);
    if (e <= EPS) { t = 0.0; s = clamp01(-c / a); }
    else {
      const double b = v3_dot(d1, d2);
      const double denom = fma(a, e, -(b * b));
      s = (denom > 0.0) ? clamp01(fma(b, f, -(c * e)) / denom) : 0.0;
      t = fma(b, s, f) / e;
      if (t < 0.0) { t = 0.0; s = clamp01(-c / a); }
      else if (t > 1.0) { t = 1.0; s = clamp01((b - c) / a); }
    }
  }
  double dv[3];
  for (int k = 0; k < 3; ++k) dv[k] = fma(s, d1[k], p1[k]) - fma(t, d2[k], p2[k]);
  return sqrt(v3_dot(dv, dv));
}

/* half derivative of the squared distance from p + s d to the box [-h, h] */
static double seg_box_g(const double* p, const double* d, const double* h, double s)
{
  double g = 0.0;
  for (int i = 0; i < 3; ++i) {
    const double c = fma(s, d[i], p[i]);
    const double e = (c > h[i]) ? c - h[i] : ((c < -h[i]) ? c + h[i] : 0.0);
    g = fma(e, d[i], g);
  }
  return g;
}

static double seg_box_f(const double* p, const double* d, const double* h, double s)
{
  double f = 0.0;
  for (int i = 0; i < 3; ++i) {
    const double c = fma(s, d[i], p[i]);
    const double e = (c > h[i]) ? c - h[i] : ((c < -h[i]) ? c + h[i] : 0.0);
    f = fma(e, e, f);
  }
  return f;
}

/* distance from segment (box frame) to the solid box with half extents h */
static double seg_box_dist(const double* p, const double* d, const double* h)
{
  double lo = 0.0, hi = 1.0;
  double glo = seg_box_g(p, d, h, 0.0), ghi = seg_box_g(p, d, h, 1.0);
  double s;
  if (glo >= 0.0) s = 0.0;
  else if (ghi <= 0.0) s = 1.0;
  else {
    for (int i = 0; i < 3; ++i) {
      if (d[i] == 0.0) continue;
      for (int sg = 0; sg < 2; ++sg) {
        const double b = ((sg ? h[i] : -h[i]) - p[i]) / d[i];
        if (!(b > lo && b < hi)) continue;
        const double gb = seg_box_g(p, d, h, b);
        if (gb <= 0.0) { lo = b; glo = gb; } else { hi = b; ghi = gb; }
      }
    }
    const double dg = ghi - glo;
    s = (dg > 0.0) ? fma(hi - lo, (-glo) / dg, lo) : lo;
  }
  return sqrt(seg_box_f(p, d, h, s));
}

/* distance from point p + s d (cylinder frame: axis z, half length hl, radius R) */
static double seg_cyl_D(const double* p, const double* d, double hl, double R, double s)
{
  const double x = fma(s, d[0], p[0]), y = fma(s, d[1], p[1]), z = fma(s, d[2], p[2]);
  const double rho = sqrt(fma(y, y, x * x));
  const double er = (rho > R) ? rho - R : 0.0;
  const double az = fabs(z);
  const double ez = (az > hl) ? az - hl : 0.0;
  return sqrt(fma(ez, ez, er * er));
}

#define GOLDEN_ITERS 48
static double seg_cyl_dist(const double* p, const double* d, double hl, double R)
{
  const double PHI = 0.6180339887498949;
  double best = seg_cyl_D(p, d, hl, R, 0.0);
  if (d[0] == 0.0 && d[1] == 0.0 && d[2] == 0.0) return best;
  const double D1 = seg_cyl_D(p, d, hl, R, 1.0);
  if (D1 < best) best = D1;
  double a = 0.0, b = 1.0;
  double x1 = b - PHI * (b - a), x2 = a + PHI * (b - a);
  double f1 = seg_cyl_D(p, d, hl, R, x1), f2 = seg_cyl_D(p, d, hl, R, x2);
  for (int it = 0; it < GOLDEN_ITERS; ++it) {
    if (f1 <= f2) { b = x2; x2 = x1; f2 = f1; x1 = b - PHI * (b - a); f1 = seg_cyl_D(p, d, hl, R, x1); }
    else { a = x1; x1 = x2; f1 = f2; x2 = a + PHI * (b - a); f2 = seg_cyl_D(p, d, hl, R, x2); }
  }
  const double Dm = seg_cyl_D(p, d, hl, R, 0.5 * (a + b));
  if (Dm < best) best = Dm;
  return best;
}

/* segment core of an SSL / SPHERE shape in world coordinates */
static void shape_segment(int type, const double* ext, const HTr* A, double* p, double* d, double* r)
{
  for (int k = 0; k < 3; ++k) p[k] = A->org[k];
  *r = ext[0];
  if (type == AFF_SHAPE_SSL) for (int k = 0; k < 3; ++k) d[k] = ext[2] * A->rot[2][k];
  else d[0] = d[1] = d[2] = 0.0;
}

static int is_seg(int type) { return type == AFF_SHAPE_SSL || type == AFF_SHAPE_SPHERE; }

/* Closest distance of two primitives; DBL_MAX for pairs without a distance
 * function (BOX-BOX, BOX-CYLINDER, CYLINDER-CYLINDER), which Rcs also leaves
 * out of its collision model when no function is registered. */
static double shape_dist(int t1, const double* e1, const HTr* A1, int t2, const double* e2, const HTr* A2)
{
  if (!is_seg(t1) && is_seg(t2)) return shape_dist(t2, e2, A2, t1, e1, A1);
  if (!is_seg(t1)) return DBL_MAX;
  double p1[3], d1[3], r1;
  shape_segment(t1, e1, A1, p1, d1, &r1);
  if (is_seg(t2)) {
    double p2[3], d2[3], r2;
    shape_segment(t2, e2, A2, p2, d2, &r2);
    return (seg_seg_dist(p1, d1, p2, d2) - r1) - r2;
  }
  /* segment in the frame of shape 2 */
  double rel[3], pl[3], dl[3];
  for (int k = 0; k < 3; ++k) rel[k] = p1[k] - A2->org[k];
  m3_mulv(pl, A2->rot, rel);
  m3_mulv(dl, A2->rot, d1);
  if (t2 == AFF_SHAPE_BOX) {
    double h[3] = {0.5 * e2[0], 0.5 * e2[1], 0.5 * e2[2]};
    return seg_box_dist(pl, dl, h) - r1;
  }
  return seg_cyl_dist(pl, dl, 0.5 * e2[2], e2[0]) - r1;
}

double oracle_shape_distance(int type1, const double* ext1, const double* A1,
                             int type2, const double* ext2, const double* A2)
{
  HTr a, b;
  htr_from12(&a, A1); htr_from12(&b, A2);
  return shape_dist(type1, ext1, &a, type2, ext2, &b);
}

/* world AABB of a shape */
static void shape_aabb(int type, const double* ext, const HTr* A, double* mn, double* mx)
{
  if (type == AFF_SHAPE_SSL || type == AFF_SHAPE_SPHERE) {
    double p[3], d[3], r;
    shape_segment(type, ext, A, p, d, &r);
    for (int k = 0; k < 3; ++k) {
      const double q = p[k] + d[k];
      mn[k] = ((p[k] < q) ? p[k] : q) - r;
      mx[k] = ((p[k] > q) ? p[k] : q) + r;
    }
  } else if (type == AFF_SHAPE_BOX) {
    for (int k = 0; k < 3; ++k) {
      const double e = fma(fabs(A->rot[2][k]), 0.5 * ext[2], fma(fabs(A->rot[1][k]), 0.5 * ext[1], fabs(A->rot[0][k]) * (0.5 * ext[0])));
      mn[k] = A->org[k] - e; mx[k] = A->org[k] + e;
    }
  } else {
    for (int k = 0; k < 3; ++k) {
      const double az = A->rot[2][k];
      const double c = 1.0 - az * az;
      const double e = fma(fabs(az), 0.5 * ext[2], ext[0] * sqrt(c > 0.0 ? c : 0.0));
      mn[k] = A->org[k] - e; mx[k] = A->org[k] + e;
    }
  }
}

/* ------------------------------------------------------- collision model */
typedef struct { int b1, b2; double dist; } Pair;

typedef struct {
  Pair* pair; int n, cap;
  double minDist; int minIdx;
  double cost;
} CollMdl;

static int body_active_shapes(const Graph* g, int b)
{
  const aff_scene_desc* d = g->d;
  for (int k = 0; k < d->body_n_shapes[b]; ++k) if (g->sdist[d->body_first_shape[b] + k]) return 1;
  return 0;
}

/* ControllerBase::computeCollisionModel: broad phase (RcsBroadPhase) + narrow
 * phase (RcsCollisionMdl_compute), restated:
 *  - a body belongs to tree T if T's root is an ancestor-or-self in the
 *    CURRENT topology (a grasped object joins the arm's tree);
 *  - pairs are enumerated over shape bodies b1 < b2; two tree bodies pair iff
 *    their trees differ; a tree body pairs with an explicit <Body> entry
 *    always, and with any other body iff their AABBs, inflated by
 *    DistanceThreshold, overlap; bodies outside all trees never pair;
 *  - pair distance = min over active shape pairs;
 *  - cost per pair with threshold thr = DistanceThreshold:
 *      d < 0: 1 - 2 d / thr;  0 <= d < thr: ((d - thr)/thr)^2;  else 0. */
static void collmdl_compute(Graph* g, CollMdl* cm)
{
  const aff_scene_desc* d = g->d;
  const double thr = d->bp_threshold;
  cm->n = 0; cm->minDist = DBL_MAX; cm->minIdx = -1; cm->cost = 0.0;
  /* tree membership + AABBs of active shape bodies */
  for (int i = 0; i < g->nSB; ++i) {
    const int b = g->sbody[i];
    g->tree_of[b] = -1;
    if (!body_active_shapes(g, b)) { g->tree_of[b] = -2; continue; }
    for (int a = b; a >= 0 && g->tree_of[b] < 0; a = g->parent[a])
      for (int t = 0; t < d->n_bp_trees; ++t) if (d->bp_tree_root[t] == a) { g->tree_of[b] = t; break; }
    double* bb = g->aabb + 6 * b;
    bb[0] = bb[1] = bb[2] = DBL_MAX; bb[3] = bb[4] = bb[5] = -DBL_MAX;
    for (int k = 0; k < d->body_n_shapes[b]; ++k) {
      const int s = d->body_first_shape[b] + k;
      if (!g->sdist[s]) continue;
      HTr Acb, A; double mn[3], mx[3];
      htr_from12(&Acb, d->shape_A_CB + 12 * s);
      htr_compose(&A, &Acb, &g->A_BI[b]);
      shape_aabb(d->shape_type[s], d->shape_extents + 3 * s, &A, mn, mx);
      for (int c = 0; c < 3; ++c) { if (mn[c] < bb[c]) bb[c] = mn[c]; if (mx[c] > bb[3 + c]) bb[3 + c] = mx[c]; }
    }
  }
  for (int i = 0; i < g->nSB; ++i) {
    const int b1 = g->sbody[i], t1 = g->tree_of[b1];
    if (t1 == -2) continue;
    for (int j = i + 1; j < g->nSB; ++j) {
      const int b2 = g->sbody[j], t2 = g->tree_of[b2];
      if (t2 == -2) continue;
      if (t1 < 0 && t2 < 0) continue;
      if (t1 >= 0 && t2 >= 0) { if (t1 == t2) continue; }
      else {
        const int other = (t1 >= 0) ? b2 : b1;
        if (!g->explicit_bp[other]) {
          const double* A = g->aabb + 6 * b1; const double* B = g->aabb + 6 * b2;
          int overlap = 1;
          for (int c = 0; c < 3; ++c) if (A[c] - thr > B[3 + c] || B[c] - thr > A[3 + c]) overlap = 0;
          if (!overlap) continue;
        }
      }
      /* narrow phase */
      double dmin = DBL_MAX;
      for (int k1 = 0; k1 < d->body_n_shapes[b1]; ++k1) {
        const int s1 = d->body_first_shape[b1] + k1;
        if (!g->sdist[s1]) continue;
        HTr Acb1, A1; htr_from12(&Acb1, d->shape_A_CB + 12 * s1); htr_compose(&A1, &Acb1, &g->A_BI[b1]);
        for (int k2 = 0; k2 < d->body_n_shapes[b2]; ++k2) {
          const int s2 = d->body_first_shape[b2] + k2;
          if (!g->sdist[s2]) continue;
          HTr Acb2, A2; htr_from12(&Acb2, d->shape_A_CB + 12 * s2); htr_compose(&A2, &Acb2, &g->A_BI[b2]);
          const double dd = shape_dist(d->shape_type[s1], d->shape_extents + 3 * s1, &A1,
                                       d->shape_type[s2], d->shape_extents + 3 * s2, &A2);
          if (dd < dmin) dmin = dd;
        }
      }
      if (cm->n == cm->cap) { cm->cap = cm->cap ? 2 * cm->cap : 256; cm->pair = (Pair*)realloc(cm->pair, sizeof(Pair) * cm->cap); }
      cm->pair[cm->n].b1 = b1; cm->pair[cm->n].b2 = b2; cm->pair[cm->n].dist = dmin;
      if (dmin < cm->minDist) { cm->minDist = dmin; cm->minIdx = cm->n; }
      if (thr > 0.0) {
        if (dmin < 0.0) cm->cost = cm->cost + (1.0 - (2.0 * dmin) / thr);
        else if (dmin < thr) { const double u = (dmin - thr) / thr; cm->cost = cm->cost + u * u; }
      }
      cm->n++;
    }
  }
}

/* ------------------------------------------------ joint-limit cost, checks */
/* ControllerBase::computeJointlimitCost / Gradient */
/* Sum of up to 64 terms in the fixed pairwise order of a 32-lane butterfly
 * (terms c and c+32 are added first, then offsets 16, 8, 4, 2, 1), so that a
 * warp reduction reproduces it bit for bit. */
static double tree_sum(const double* v, int n)
{
  double a[32];
  for (int i = 0; i < 32; ++i) a[i] = (i < n ? v[i] : 0.0) + (i + 32 < n ? v[i + 32] : 0.0);
  for (int off = 16; off >= 1; off >>= 1)
    for (int i = 0; i < off; ++i) a[i] = a[i] + a[i + off];
  return a[0];
}

static double jl_cost(const Graph* g)
{
  const aff_scene_desc* d = g->d;
  double term[64];
  for (int c = 0; c < g->nJ && c < 64; ++c) {
    const int j = g->ikj[c];
    const double dq = g->q[j] - d->joint_q0[j];
    const double r = (dq > 0.0) ? d->joint_q_max[j] - d->joint_q0[j] : d->joint_q0[j] - d->joint_q_min[j];
    const double u = dq / r;
    term[c] = (0.5 * d->joint_weight_jl[j]) * (u * u);
  }
  return tree_sum(term, g->nJ < 64 ? g->nJ : 64);
}

static void jl_gradient(const Graph* g, double* dH)
{
  const aff_scene_desc* d = g->d;
  for (int c = 0; c < g->nJ; ++c) {
    const int j = g->ikj[c];
    const double dq = g->q[j] - d->joint_q0[j];
    const double r = (dq > 0.0) ? d->joint_q_max[j] - d->joint_q0[j] : d->joint_q0[j] - d->joint_q_min[j];
    dH[c] = (d->joint_weight_jl[j] * dq) / (r * r);
  }
}

/* RcsGraph_checkJointSpeeds over the IK joints (constrained dofs never move) */
static double check_joint_speeds(const Graph* g, const double* dq, double dt)
{
  const aff_scene_desc* d = g->d;
  double scale = 1.0;
  for (int c = 0; c < g->nJ; ++c) {
    const int j = g->ikj[c];
    const double lim = d->joint_speed_limit[j];
    if (lim >= NO_LIMIT) continue;
    const double a = fabs(dq[c]);
    const double mx = lim * dt;
    if (a > mx) { const double s = mx / a; if (s < scale) scale = s; }
  }
  return scale;
}

static int num_joint_limits_violated(const Graph* g)
{
  const aff_scene_desc* d = g->d;
  int n = 0;
  for (int c = 0; c < g->nJ; ++c) {
    const int j = g->ikj[c];
    if (g->q[j] < d->joint_q_min[j] || g->q[j] > d->joint_q_max[j]) ++n;
  }
  return n;
}

/* TrajectoryPredictor::checkState (src/TrajectoryPredictor.cpp:756-886) */
static int check_state(const Graph* g, const CollMdl* cm, const double* qdot_ik, int* id0, int* id1)
{
  if (check_joint_speeds(g, qdot_ik, 1.0) < 1.0) { *id0 = -1; *id1 = -1; return AFF_CODE_SPEED; }
  const int aor = num_joint_limits_violated(g);
  if (aor > 0) { *id0 = aor; *id1 = -1; return AFF_CODE_JOINT_LIMIT; }
  if (cm->minIdx != -1 && cm->minDist < 0.001) {
    *id0 = cm->pair[cm->minIdx].b1; *id1 = cm->pair[cm->minIdx].b2;
    return AFF_CODE_COLLISION;
  }
  return AFF_CODE_OK;
}

/* relative position component i of `body` w.r.t. base, in base coordinates */
static double rel_comp(const Graph* g, int base, int body, int i)
{
  double dlt[3];
  for (int k = 0; k < 3; ++k) dlt[k] = g->A_BI[body].org[k] - g->A_BI[base].org[k];
  const double* R = g->A_BI[base].rot[i];
  return fma(R[2], dlt[2], fma(R[1], dlt[1], R[0] * dlt[0]));
}

/* RcsGraph_1dPosJacobian(graph, eff, base, NULL, idx, J) scaled by f, added to dH */
static void add_1d_pos_jac(const Graph* g, int eff, int base, int idx, double f, double* dH)
{
  aff_task_desc t;
  memset(&t, 0, sizeof(t));
  t.kind = AFF_TASK_X + idx; t.effector = eff; t.ref_body = base; t.ref_frame = -2; /* refFrame NULL = world */
  double* J = (double*)malloc(sizeof(double) * (size_t)(g->nJ + 1));
  task_J(g, &t, J);
  for (int c = 0; c < g->nJ; ++c) dH[c] = dH[c] + J[c] * f;
  free(J);
}

/* TrajectoryPredictor::addElbowNullspace (src/TrajectoryPredictor.cpp:448-525) */
static void add_elbow_nullspace(const Graph* g, double* dH)
{
  const aff_scene_desc* d = g->d;
  const double gain = 5.0, clipv = -0.5;
  const int base = d->ns_base;
  if (base < 0) return;
  for (int pass = 0; pass < 2; ++pass) {
    const double boundary = pass ? 0.0 : 0.15;
    for (int side = 0; side < 2; ++side) {
      const int el = pass ? d->ns_hand[side] : d->ns_forearm[side];
      const int sh = d->ns_upperarm[side];
      if (el < 0 || sh < 0) continue;
      const double yS = rel_comp(g, base, sh, 1), yE = rel_comp(g, base, el, 1);
      if (side == 0) {
        const double dist = clipd((yE - yS) - boundary, clipv, 0.0);
        add_1d_pos_jac(g, el, base, 1, dist * gain, dH);
      } else {
        const double dist = clipd((-yE + yS) - boundary, clipv, 0.0);
        add_1d_pos_jac(g, el, base, 1, -dist * gain, dH);
      }
    }
  }
}

/* TrajectoryPredictor::addWristNullspace (src/TrajectoryPredictor.cpp:527-570) */
static void add_wrist_nullspace(const Graph* g, double* dH)
{
  const aff_scene_desc* d = g->d;
  const double boundary = 0.25, clipv = -0.5, gain = 5.0;
  const int base = d->ns_base;
  if (base < 0) return;
  for (int side = 0; side < 2; ++side) {
    const int wr = d->ns_hand[side];
    if (wr < 0) continue;
    const double x = rel_comp(g, base, wr, 0), z = rel_comp(g, base, wr, 2);
    const double dBound = (z > 1.0) ? 0.4 * (z - 1.0) : 0.0;
    const double dist = clipd(x - (boundary + dBound), clipv, 0.0);
    add_1d_pos_jac(g, wr, base, 0, dist * gain, dH);
  }
}

/* ------------------------------------------------ trajectory controller */
typedef struct { double t, x, xd, xdd; int flag; int alive; } Via;
typedef struct { double t, horizon; int on; int done; } ActPt;

typedef struct {
  int nTasks, xdim;
  const aff_task_desc* tasks;
  int* toff;            /* [nTasks+1] offset of each task in the stacked x */
  double* st;           /* [xdim*3]   (x, xd, xdd) per dimension            */
  Via** via; int* nvia; /* per dimension, sorted by time                    */
  ActPt** act; int* nact; /* per task, sorted by time                       */
  int* active;          /* per task                                         */
  double* lastSwitchT; double* lastSwitchH; /* per task                      */
  const aff_constraint* gcon; int ngcon; unsigned char* gfired; /* graph constraints */
  double tnow, endAbs, startAbs;
} TrajCtrl;

static int cmp_via(const void* a, const void* b)
{
  const Via* x = (const Via*)a; const Via* y = (const Via*)b;
  if (x->t < y->t) return -1;
  if (x->t > y->t) return 1;
  return x->alive - y->alive; /* alive doubles as insertion order during the sort */
}

static int cmp_act(const void* a, const void* b)
{
  const ActPt* x = (const ActPt*)a; const ActPt* y = (const ActPt*)b;
  if (x->t < y->t) return -1;
  if (x->t > y->t) return 1;
  return x->done - y->done;
}

static int tc_init(TrajCtrl* tc, const aff_task_desc* tasks, int nTasks, const aff_constraint* cons, int nCons)
{
  memset(tc, 0, sizeof(*tc));
  tc->nTasks = nTasks; tc->tasks = tasks;
  tc->toff = (int*)malloc(sizeof(int) * (nTasks + 1));
  tc->xdim = 0;
  for (int i = 0; i < nTasks; ++i) { const int dm = task_dim(&tasks[i]); if (dm < 0) return -1; tc->toff[i] = tc->xdim; tc->xdim += dm; }
  tc->toff[nTasks] = tc->xdim;
  tc->st = (double*)calloc((size_t)tc->xdim * 3 + 1, sizeof(double));
  tc->via = (Via**)calloc(tc->xdim + 1, sizeof(Via*));
  tc->nvia = (int*)calloc(tc->xdim + 1, sizeof(int));
  tc->act = (ActPt**)calloc(nTasks + 1, sizeof(ActPt*));
  tc->nact = (int*)calloc(nTasks + 1, sizeof(int));
  tc->active = (int*)calloc(nTasks + 1, sizeof(int));
  tc->lastSwitchT = (double*)calloc(nTasks + 1, sizeof(double));
  tc->lastSwitchH = (double*)calloc(nTasks + 1, sizeof(double));
  tc->gcon = cons; tc->ngcon = nCons;
  tc->gfired = (unsigned char*)calloc(nCons + 1, 1);
  for (int i = 0; i < tc->xdim; ++i) tc->via[i] = (Via*)calloc(nCons + 1, sizeof(Via));
  for (int i = 0; i < nTasks; ++i) { tc->act[i] = (ActPt*)calloc(nCons + 1, sizeof(ActPt)); tc->lastSwitchH[i] = -1.0; }
  double endT = 0.0, startT = DBL_MAX;
  for (int k = 0; k < nCons; ++k) {
    const aff_constraint* c = &cons[k];
    if (c->t > endT) endT = c->t;
    if (c->t < startT) startT = c->t;
    if (c->kind == AFF_CON_VIA) {
      if (c->task < 0 || c->task >= nTasks || c->dim < 0 || c->dim >= task_dim(&tasks[c->task])) return -1;
      const int dIdx = tc->toff[c->task] + c->dim;
      Via* v = &tc->via[dIdx][tc->nvia[dIdx]];
      v->t = c->t; v->x = c->x; v->xd = c->xd; v->xdd = c->xdd; v->flag = c->flag & 7; v->alive = tc->nvia[dIdx];
      tc->nvia[dIdx]++;
    } else if (c->kind == AFF_CON_ACTIVATION) {
      if (c->task < 0 || c->task >= nTasks) return -1;
      ActPt* a = &tc->act[c->task][tc->nact[c->task]];
      a->t = c->t; a->horizon = c->horizon; a->on = c->flag ? 1 : 0; a->done = tc->nact[c->task];
      tc->nact[c->task]++;
    }
  }
  for (int i = 0; i < tc->xdim; ++i) {
    qsort(tc->via[i], tc->nvia[i], sizeof(Via), cmp_via);
    for (int k = 0; k < tc->nvia[i]; ++k) tc->via[i][k].alive = 1;
  }
  for (int i = 0; i < nTasks; ++i) {
    qsort(tc->act[i], tc->nact[i], sizeof(ActPt), cmp_act);
    for (int k = 0; k < tc->nact[i]; ++k) tc->act[i][k].done = 0;
  }
  tc->endAbs = endT;
  tc->startAbs = (startT == DBL_MAX) ? 0.0 : (startT < 0.0 ? 0.0 : startT);
  return 0;
}

static void tc_free(TrajCtrl* tc)
{
  for (int i = 0; i < tc->xdim; ++i) free(tc->via[i]);
  for (int i = 0; i < tc->nTasks; ++i) free(tc->act[i]);
  free(tc->toff); free(tc->st); free(tc->via); free(tc->nvia); free(tc->act); free(tc->nact);
  free(tc->active); free(tc->lastSwitchT); free(tc->lastSwitchH); free(tc->gfired);
}

/* TrajectoryControllerBase::step(dt), restated:
 *  1. trajectories of inactive tasks are re-initialised at the current task
 *     value with zero velocity / acceleration;
 *  2. every 1-D trajectory advances by dt through its receding window;
 *  3. time advances; constraints whose time has been reached fire:
 *     ConnectBodyConstraint, CollisionModelConstraint, activation switches;
 *     passed via points are dropped.
 * Returns the remaining time to the last constraint. */
static double tc_step(TrajCtrl* tc, Graph* g, double dt)
{
  for (int i = 0; i < tc->nTasks; ++i) {
    if (tc->active[i]) continue;
    double x[AFF_MAX_TASK_JOINTS];
    task_x(g, &tc->tasks[i], x);
    const int dm = tc->toff[i + 1] - tc->toff[i];
    for (int k = 0; k < dm; ++k) { double* s = tc->st + 3 * (tc->toff[i] + k); s[0] = x[k]; s[1] = 0.0; s[2] = 0.0; }
  }
  for (int dI = 0; dI < tc->xdim; ++dI) {
    double rows[5 * 64];
    int n = 0;
    for (int k = 0; k < tc->nvia[dI] && n < 64; ++k) {
      const Via* v = &tc->via[dI][k];
      if (!v->alive) continue;
      rows[5 * n] = v->t - tc->tnow; rows[5 * n + 1] = v->x; rows[5 * n + 2] = v->xd; rows[5 * n + 3] = v->xdd; rows[5 * n + 4] = (double)v->flag;
      ++n;
    }
    oracle_via_step(tc->st + 3 * dI, rows, n, TRAJ_HORIZON, dt);
  }
  tc->tnow = tc->tnow + dt;
  /* graph constraints */
  for (int k = 0; k < tc->ngcon; ++k) {
    const aff_constraint* c = &tc->gcon[k];
    if (tc->gfired[k]) continue;
    if (c->kind == AFF_CON_CONNECT) {
      if (c->t - tc->tnow < TRAJ_EPS) { graph_connect(g, c->body_a, c->body_b, c->dim == 1); tc->gfired[k] = 1; }
    } else if (c->kind == AFF_CON_COLLISION) {
      /* toggleTime -= dt; fires when toggleTime < 0 (CollisionModelConstraint.h:101-105) */
      if (c->t - tc->tnow < 0.0) {
        const aff_scene_desc* d = g->d;
        const int b = c->body_a;
        if (b >= 0 && b < g->nB)
          for (int s = 0; s < d->body_n_shapes[b]; ++s) g->sdist[d->body_first_shape[b] + s] = c->flag ? 1 : 0;
        tc->gfired[k] = 1;
      }
    }
  }
  for (int i = 0; i < tc->nTasks; ++i)
    for (int k = 0; k < tc->nact[i]; ++k) {
      ActPt* a = &tc->act[i][k];
      if (a->done) continue;
      if (a->t - tc->tnow < TRAJ_EPS) { tc->active[i] = a->on; tc->lastSwitchT[i] = a->t; tc->lastSwitchH[i] = a->horizon; a->done = 1; }
      else break;
    }
  for (int dI = 0; dI < tc->xdim; ++dI)
    for (int k = 0; k < tc->nvia[dI]; ++k) {
      Via* v = &tc->via[dI][k];
      if (v->alive && v->t - tc->tnow < TRAJ_EPS) v->alive = 0;
    }
  const double rem = tc->endAbs - tc->tnow;
  return rem > 0.0 ? rem : 0.0;
}

/* TrajectoryControllerBase::computeBlending: min over tasks of the ramp that
 * goes to 0 at every activation switch and back to 1 over its horizon. */
static double tc_blending(const TrajCtrl* tc)
{
  double bl = 1.0;
  for (int i = 0; i < tc->nTasks; ++i) {
    double b = 1.0;
    for (int k = 0; k < tc->nact[i]; ++k) {
      const ActPt* a = &tc->act[i][k];
      if (a->done) continue;
      if (a->horizon > 0.0) b = clipd((a->t - tc->tnow) / a->horizon, 0.0, 1.0);
      break;
    }
    if (tc->lastSwitchH[i] > 0.0) {
      const double up = clipd((tc->tnow - tc->lastSwitchT[i]) / tc->lastSwitchH[i], 0.0, 1.0);
      if (up < b) b = up;
    }
    if (b < bl) bl = b;
  }
  return bl;
}

/* ------------------------------------------------------------ the predictor */
int oracle_num_steps(const aff_constraint* cons, const aff_candidate* cand, const aff_opts* opts)
{
  double endT = 0.0;
  for (int k = 0; k < cand->n_constraints; ++k) { const double t = cons[cand->first_constraint + k].t; if (t > endT) endT = t; }
  return (int)lround(endT / opts->dt) + opts->n_after_steps;
}

typedef struct {
  Graph g; TrajCtrl tc; CollMdl cm;
  const aff_task_desc* tasks; int nTasks;
  double* J; double* dH; double* dq; double* w; double* qdot_ik;
  unsigned char* emask; unsigned char* amask; int* jmask;
} Pred;

/* eMask of computeIK (src/TrajectoryPredictor.cpp:646-664): zero for every IK
 * joint that moves a body owning rigid-body joints. */
static void compute_emask(Pred* P)
{
  Graph* g = &P->g;
  for (int c = 0; c < g->nJ; ++c) P->emask[c] = 1;
  for (int b = 0; b < g->nB; ++b) {
    if (!g->d->body_rbj[b]) continue;
    /* RcsGraph_computeJointRecursionMask: joints from the body up to the root */
    for (int a = b; a >= 0; a = g->parent[a])
      for (int k = 0; k < g->d->body_n_joints[a]; ++k) {
        const int c = g->jacobi[g->d->body_first_joint[a] + k];
        if (c >= 0) P->emask[c] = 0;
      }
  }
}

/* TrajectoryPredictor::computeIK (src/TrajectoryPredictor.cpp:573-753) */
static int compute_ik(Pred* P, const int* a_des, const double* x_des, double dt, double alpha,
                      double lambda, double qFilt, double phase, int* id0, int* id1)
{
  Graph* g = &P->g;
  const aff_scene_desc* d = g->d;
  const int nJ = g->nJ;
  /* dx_des = computeDX(x) and J for the active tasks, compressed */
  double dx[64];
  int m = 0;
  for (int i = 0; i < P->nTasks; ++i) {
    if (!a_des[i]) continue;
    const int dm = task_dim(&P->tasks[i]);
    task_dx(g, &P->tasks[i], x_des + P->tc.toff[i], dx + m);
    task_J(g, &P->tasks[i], P->J + (size_t)m * nJ);
    m += dm;
  }
  /* dH = alpha * joint limit gradient */
  jl_gradient(g, P->dH);
  for (int c = 0; c < nJ; ++c) P->dH[c] = P->dH[c] * alpha;
  /* extra null space: elbows and wrists (:596-625) */
  {
    double* ex = (double*)calloc(nJ + 1, sizeof(double));
    double* tmp = (double*)calloc(nJ + 1, sizeof(double));
    add_elbow_nullspace(g, tmp);
    for (int c = 0; c < nJ; ++c) ex[c] = ex[c] + tmp[c];
    memset(tmp, 0, sizeof(double) * nJ);
    add_wrist_nullspace(g, tmp);
    for (int c = 0; c < nJ; ++c) ex[c] = (ex[c] + tmp[c]) * alpha;
    const double scale = check_joint_speeds(g, ex, dt);
    if (scale < 1.0) for (int c = 0; c < nJ; ++c) ex[c] = ex[c] * (0.99999 * scale);
    for (int c = 0; c < nJ; ++c) P->dH[c] = P->dH[c] + ex[c] * phase;
    free(ex); free(tmp);
  }
  /* joint masks (:636-675) */
  for (int c = 0; c < nJ; ++c) {
    int on = 0;
    for (int r = 0; r < m; ++r) if (P->J[(size_t)r * nJ + c] != 0.0) on = 1;
    P->amask[c] = (unsigned char)on;
    P->jmask[c] += on;
    if (!(on || P->emask[c])) P->dH[c] = P->dH[c] * 0.0;
  }
  /* right inverse (:683-693) */
  if (!oracle_right_inverse(m, nJ, P->J, P->w, dx, P->dH, lambda, P->dq)) return AFF_CODE_SINGULAR;
  /* speed and acceleration limits (:696-723) */
  {
    const double scale = check_joint_speeds(g, P->dq, dt);
    if (g_trace > 0 && scale < 1.0) { int jm = 0; for (int c = 0; c < nJ; ++c) if (fabs(P->dq[c]) / d->joint_speed_limit[g->ikj[c]] > fabs(P->dq[jm]) / d->joint_speed_limit[g->ikj[jm]]) jm = c; fprintf(stderr, "   speed scale %.3f joint %d m %d\n", scale, jm, m); }
    if (scale < 1.0) for (int c = 0; c < nJ; ++c) P->dq[c] = P->dq[c] * (0.99999 * scale);
    const double dynLim = -0.99 * qFilt + 1.0;
    const double invdt = 1.0 / dt;
    for (int c = 0; c < nJ; ++c) {
      const int j = g->ikj[c];
      double qd = P->dq[c] * invdt;
      const double acc = d->joint_acc_limit[j];
      if (acc < NO_LIMIT) {
        const double lim = (dynLim * acc) * dt;
        const double dv = qd - P->qdot_ik[c];
        if (dv > lim) qd = P->qdot_ik[c] + lim;
        else if (dv < -lim) qd = P->qdot_ik[c] - lim;
      }
      P->qdot_ik[c] = qd;
    }
  }
  /* integration and forward kinematics (:730-733) */
  for (int c = 0; c < nJ; ++c) { const int j = g->ikj[c]; g->q[j] = g->q[j] + P->dq[c]; g->qdot[j] = P->qdot_ik[c]; }
  graph_fk(g);
  collmdl_compute(g, &P->cm);
  return check_state(g, &P->cm, P->qdot_ik, id0, id1);
}

int oracle_predict(const aff_scene_desc* scene, const aff_task_desc* tasks_all, const aff_constraint* cons_all,
                   const aff_candidate* cand, const aff_opts* opts, aff_result* out, double* q_log, size_t q_log_rows)
{
  Pred P;
  if (g_trace < 0) { const char* e = getenv("AFF_ORACLE_TRACE"); g_trace = e ? atoi(e) : 0; }
  memset(&P, 0, sizeof(P));
  memset(out, 0, sizeof(*out));
  out->first_fail_step = -1; out->fail_step = -1; out->fail_id0 = -1; out->fail_id1 = -1;
  out->min_dist_b1 = -1; out->min_dist_b2 = -1; out->min_dist = DBL_MAX;
  if (graph_init(&P.g, scene, cand) != 0) { graph_free(&P.g); return -1; }
  Graph* g = &P.g;
  P.tasks = tasks_all + cand->first_task; P.nTasks = cand->n_tasks;
  if (tc_init(&P.tc, P.tasks, P.nTasks, cons_all + cand->first_constraint, cand->n_constraints) != 0) {
    tc_free(&P.tc); graph_free(g); return -1;
  }
  const int nJ = g->nJ;
  const double dt = opts->dt;
  P.J = (double*)calloc((size_t)(P.tc.xdim + 1) * nJ, sizeof(double));
  P.dH = (double*)calloc(nJ + 1, sizeof(double));
  P.dq = (double*)calloc(nJ + 1, sizeof(double));
  P.w = (double*)calloc(nJ + 1, sizeof(double));
  P.qdot_ik = (double*)calloc(nJ + 1, sizeof(double));
  P.emask = (unsigned char*)calloc(nJ + 1, 1);
  P.amask = (unsigned char*)calloc(nJ + 1, 1);
  P.jmask = (int*)calloc(nJ + 1, sizeof(int));
  for (int c = 0; c < nJ; ++c) P.w[c] = scene->joint_weight_metric[g->ikj[c]];

  /* src/TrajectoryPredictor.cpp:92-103 */
  double endTime = P.tc.endAbs;
  const double motionDuration = P.tc.endAbs - P.tc.startAbs;
  const int nSteps = (int)lround(endTime / dt) + opts->n_after_steps;
  double t = 0.0, qFilt = 0.0, err = 0.0;
  int rows = 0;

  graph_fk(g);
  if (q_log && q_log_rows > 0) for (int c = 0; c < nJ; ++c) q_log[c] = g->q[g->ikj[c]];
  rows = 1;                                     /* step-0 transform row, :111-118 */
  collmdl_compute(g, &P.cm);                    /* :121 */
  out->min_dist = P.cm.minDist;                 /* :128-130 */
  if (P.cm.minIdx >= 0) { out->min_dist_b1 = P.cm.pair[P.cm.minIdx].b1; out->min_dist_b2 = P.cm.pair[P.cm.minIdx].b2; }

  /* check(): checkLimits(jl, coll, speed, 0.0, 0.0, 0.01), :135-140, :426-441 */
  int ok0 = 1;
  if (num_joint_limits_violated(g) > 0) ok0 = 0;
  if (P.cm.minIdx >= 0 && P.cm.minDist < 0.01) ok0 = 0;
  if (check_joint_speeds(g, P.qdot_ik, 1.0) < 1.0) ok0 = 0;
  if (!ok0) {
    out->success = 0; out->code = AFF_CODE_INITIAL_STATE; out->first_code = AFF_CODE_INITIAL_STATE;
    out->first_fail_step = 0; out->fail_step = 0; out->n_steps = 0;
    goto done;
  }

  {
    int* a_des = (int*)calloc(P.nTasks + 1, sizeof(int));
    int* a_prev = (int*)calloc(P.nTasks + 1, sizeof(int));
    double* x_des = (double*)calloc(P.tc.xdim + 1, sizeof(double));
    const double alpha = opts->alpha, lambda = opts->lambda, qFiltDecay = opts->q_filt_decay;
    int success = 1;
    compute_emask(&P);
    for (int iter = 0; iter < nSteps; ++iter) {
      const int successPrev = success;
      /* :166-180 */
      int taskSwitch = 0;
      for (int i = 0; i < P.nTasks; ++i) if (a_prev[i] != a_des[i]) taskSwitch = 1;
      if (taskSwitch) qFilt = 1.0;
      if (qFiltDecay <= 0.0) qFilt = 0.0;
      else qFilt = clipd(qFilt - (1.0 / qFiltDecay) * dt, 0.0, 1.0);
      /* :182-188 */
      endTime = tc_step(&P.tc, g, dt);
      compute_emask(&P);   /* topology may have changed (ConnectBodyConstraint) */
      for (int i = 0; i < P.tc.xdim; ++i) x_des[i] = P.tc.st[3 * i];
      int anyActive = 0;
      for (int i = 0; i < P.nTasks; ++i) { a_prev[i] = a_des[i]; a_des[i] = P.tc.active[i]; if (a_des[i]) anyActive = 1; }
      double blending = tc_blending(&P.tc);
      if (!anyActive) blending = 0.0;               /* :192-195 */
      /* :202-203 */
      const double phase = (motionDuration > 0.0) ? 1.0 - (endTime / motionDuration) : 0.0;
      const double phaseScale = o_sin(AFF_PI * phase);
      int id0 = -1, id1 = -1;
      const int ikRes = compute_ik(&P, a_des, x_des, dt, blending * alpha, lambda, qFilt, phaseScale, &id0, &id1);
      if (ikRes != 0) {                             /* :220-224 */
        if (success) { out->first_code = ikRes; out->first_fail_step = iter; }
        success = 0; out->code = ikRes; out->fail_step = iter; out->fail_id0 = id0; out->fail_id1 = id1;
      }
      out->jl_cost = out->jl_cost + jl_cost(g);     /* :226-227 */
      out->coll_cost = out->coll_cost + P.cm.cost;
      if (success) {                                /* :297-337 */
        double emax = 0.0; int etask = -1, edim = -1;
        for (int i = 0; i < P.nTasks; ++i) {
          if (!a_des[i]) continue;
          double dxe[AFF_MAX_TASK_JOINTS];
          const int dm = task_dim(&P.tasks[i]);
          task_dx(g, &P.tasks[i], x_des + P.tc.toff[i], dxe);
          for (int k = 0; k < dm; ++k) { const double a = fabs(dxe[k]); if (a > emax) { emax = a; etask = i; edim = k; } }
        }
        if (emax > err) err = emax;
        if (err > MAX_TRACKING_ERROR) {
          success = 0; out->code = AFF_CODE_TRACKING; out->first_code = AFF_CODE_TRACKING;
          out->first_fail_step = iter; out->fail_step = iter; out->fail_id0 = etask; out->fail_id1 = edim;
        }
      }
      if (g_trace > 0 && (iter % g_trace == 0)) {
        fprintf(stderr, "it %d t %.2f blend %.3f phase %.3f ik %d a[", iter, P.tc.tnow, blending, phaseScale, ikRes);
        for (int i = 0; i < P.nTasks; ++i) fprintf(stderr, "%d", a_des[i]);
        fprintf(stderr, "] ");
        for (int i = 0; i < P.nTasks; ++i) {
          double xc[AFF_MAX_TASK_JOINTS], dxe[AFF_MAX_TASK_JOINTS]; const int dm = task_dim(&P.tasks[i]);
          task_x(g, &P.tasks[i], xc); task_dx(g, &P.tasks[i], x_des + P.tc.toff[i], dxe);
          fprintf(stderr, " T%d:", i);
          for (int k = 0; k < dm; ++k) fprintf(stderr, "(%.3f>%.3f|%.1e)", xc[k], x_des[P.tc.toff[i] + k], dxe[k]);
        }
        fprintf(stderr, " minD %.4f\n", P.cm.minDist);
      }
      if (P.cm.minDist < out->min_dist) {           /* :346-353 */
        out->min_dist = P.cm.minDist;
        out->min_dist_b1 = P.cm.pair[P.cm.minIdx].b1; out->min_dist_b2 = P.cm.pair[P.cm.minIdx].b2;
      }
      if (q_log && (size_t)rows < q_log_rows) for (int c = 0; c < nJ; ++c) q_log[(size_t)rows * nJ + c] = g->q[g->ikj[c]];
      rows++;
      /* :365-373, the reference's commented-out `break`: opt-in only */
      if (opts->early_exit && !success && successPrev) break;
      t += dt;
    }
    out->jl_cost = out->jl_cost / (double)rows;     /* :383-384 */
    out->coll_cost = out->coll_cost / (double)rows;
    out->success = success;
    out->n_steps = rows;
    out->track_err = err;
    for (int c = 0; c < nJ && c < 64; ++c) if (P.jmask[c] > 0) out->jmask_bits |= (1ull << c);
    free(a_des); free(a_prev); free(x_des);
  }
done:
  free(P.J); free(P.dH); free(P.dq); free(P.w); free(P.qdot_ik); free(P.emask); free(P.amask); free(P.jmask);
  free(P.cm.pair);
  tc_free(&P.tc);
  graph_free(g);
  return 0;
}

/* ------------------------------------------------------- host thread pool */
typedef struct {
  const aff_scene_desc* scene; const aff_task_desc* tasks; const aff_constraint* cons;
  const aff_candidate* cands; size_t n; const aff_opts* opts; aff_result* out;
  size_t next; pthread_mutex_t mtx; int err;
} PoolJob;

static void* pool_worker(void* arg)
{
  PoolJob* job = (PoolJob*)arg;
  for (;;) {
    pthread_mutex_lock(&job->mtx);
    const size_t i = job->next++;
    pthread_mutex_unlock(&job->mtx);
    if (i >= job->n) break;
    if (oracle_predict(job->scene, job->tasks, job->cons, &job->cands[i], job->opts, &job->out[i], NULL, 0) != 0) job->err = 1;
    job->out[i].idx = (int32_t)i;   /* idx assigned after predict returns (ActionComponent.cpp:191) */
  }
  return NULL;
}

int oracle_predict_batch(const aff_scene_desc* scene, const aff_task_desc* tasks, const aff_constraint* cons,
                         const aff_candidate* cands, size_t n_cands, const aff_opts* opts, int n_threads, aff_result* out)
{
  PoolJob job;
  memset(&job, 0, sizeof(job));
  job.scene = scene; job.tasks = tasks; job.cons = cons; job.cands = cands; job.n = n_cands; job.opts = opts; job.out = out;
  pthread_mutex_init(&job.mtx, NULL);
  if (n_threads < 1) n_threads = 1;
  if ((size_t)n_threads > n_cands) n_threads = (int)(n_cands ? n_cands : 1);   /* min(hw, nSolutions) */
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
  for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, pool_worker, &job);
  for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
  free(th);
  pthread_mutex_destroy(&job.mtx);
  return job.err ? -1 : 0;
}

/* ------------------------------------------------- test-only entry points */
int oracle_fk(const aff_scene_desc* scene, const aff_candidate* cand, double* A_BI)
{
  Graph g;
  if (graph_init(&g, scene, cand) != 0) { graph_free(&g); return -1; }
  graph_fk(&g);
  for (int b = 0; b < g.nB; ++b) htr_to12(&g.A_BI[b], A_BI + 12 * b);
  graph_free(&g);
  return 0;
}

static int graph_for_q(Graph* g, const aff_scene_desc* scene, const double* q_ik)
{
  if (graph_init(g, scene, NULL) != 0) return -1;
  if (q_ik) for (int c = 0; c < g->nJ; ++c) g->q[g->ikj[c]] = q_ik[c];
  graph_fk(g);
  return 0;
}

int oracle_task_eval(const aff_scene_desc* scene, const aff_task_desc* task, const double* q_ik, double* x, double* J)
{
  Graph g;
  if (graph_for_q(&g, scene, q_ik) != 0) { graph_free(&g); return -1; }
  const int dm = task_dim(task);
  if (dm > 0) { if (x) task_x(&g, task, x); if (J) task_J(&g, task, J); }
  graph_free(&g);
  return dm;
}

int oracle_task_dx(const aff_scene_desc* scene, const aff_task_desc* task, const double* q_ik, const double* x_des, double* dx)
{
  Graph g;
  if (graph_for_q(&g, scene, q_ik) != 0) { graph_free(&g); return -1; }
  const int dm = task_dim(task);
  if (dm > 0) task_dx(&g, task, x_des, dx);
  graph_free(&g);
  return dm;
}

#include "Graph.h"

#include <algorithm>
#include <cctype>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace affgpu
{

HTr HTr::identity()
{
  HTr A;
  std::memset(&A, 0, sizeof(A));
  A.rot[0][0] = A.rot[1][1] = A.rot[2][2] = 1.0;
  return A;
}

HTr HTr::from12(const double* v)
{
  HTr A;
  for (int i = 0; i < 3; ++i) A.org[i] = v[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A.rot[i][j] = v[3 + 3 * i + j];
  return A;
}

void HTr::to12(double* v) const
{
  for (int i = 0; i < 3; ++i) v[i] = org[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[3 + 3 * i + j] = rot[i][j];
}

HTr compose(const HTr& T, const HTr& P)
{
  HTr r;
  for (int i = 0; i < 3; ++i)
    r.org[i] = P.org[i] + P.rot[0][i] * T.org[0] + P.rot[1][i] * T.org[1] + P.rot[2][i] * T.org[2];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      r.rot[i][j] = T.rot[i][0] * P.rot[0][j] + T.rot[i][1] * P.rot[1][j] + T.rot[i][2] * P.rot[2][j];
  return r;
}

static void read12(std::istringstream& is, HTr& A)
{
  double v[12];
  for (double& x : v) is >> x;
  A = HTr::from12(v);
}

Graph Graph::load(const std::string& file, std::vector<std::string>* tail)
{
  std::ifstream in(file);
  if (!in) throw std::runtime_error("cannot open scene file " + file);
  Graph g;
  std::string line, tok;
  if (!std::getline(in, line) || line.rfind("AFFSCENE", 0) != 0) throw std::runtime_error(file + ": not an .affscene file");
  std::vector<std::string> treeNames, bpNames;
  bool inTail = false;
  while (std::getline(in, line)) {
    if (inTail) { if (tail) tail->push_back(line); continue; }
    std::istringstream is(line);
    is >> tok;
    if (tok == "graph") { is >> g.name; }
    else if (tok == "counts") { size_t nb, nj, ns; is >> nb >> nj >> ns; g.bodies.reserve(nb); g.joints.reserve(nj); g.shapes.reserve(ns); }
    else if (tok == "body") {
      Body b; int id, rbj, nj, ns;
      is >> id >> b.name >> b.parent >> rbj >> nj >> ns;
      read12(is, b.A_BP);
      b.rigid_body_joints = rbj != 0;
      b.firstJoint = (int)g.joints.size(); b.firstShape = (int)g.shapes.size();
      b.A_BI = HTr::identity();
      if (id != (int)g.bodies.size()) throw std::runtime_error(file + ": body ids out of order");
      g.bodyIndex_[b.name] = id;
      g.bodies.push_back(b);
    }
    else if (tok == "joint") {
      Joint j; int con;
      is >> j.name >> j.type >> con >> j.q_init >> j.q0 >> j.q_min >> j.q_max >> j.speedLimit >> j.accLimit >> j.weightJL >> j.weightMetric;
      read12(is, j.A_JP);
      j.constrained = con != 0;
      j.body = (int)g.bodies.size() - 1;
      j.jacobiIndex = j.constrained ? -1 : g.nJ++;
      g.jointIndex_[j.name] = (int)g.joints.size();
      g.joints.push_back(j);
      g.bodies.back().nJoints++;
    }
    else if (tok == "shape") {
      Shape s; int dist;
      is >> s.type >> dist >> s.extents[0] >> s.extents[1] >> s.extents[2];
      read12(is, s.A_CB);
      s.distance = dist != 0;
      s.body = (int)g.bodies.size() - 1;
      g.shapes.push_back(s);
      g.bodies.back().nShapes++;
    }
    else if (tok == "broadphase") { size_t nt, nb; is >> g.bpThreshold >> nt >> nb; }
    else if (tok == "tree") { std::string n; is >> n; treeNames.push_back(n); }
    else if (tok == "bpbody") { std::string n; is >> n; bpNames.push_back(n); }
    else if (tok == "entities") { inTail = true; if (tail) tail->push_back(line); }
    if (!inTail && is.fail() && tok != "graph") throw std::runtime_error(file + ": parse error in line: " + line.substr(0, 60));
  }
  for (const auto& n : treeNames) { int b = g.getBodyByName(n); if (b < 0) throw std::runtime_error("broad phase tree root not found: " + n); g.bpTrees.push_back(b); }
  for (const auto& n : bpNames) { int b = g.getBodyByName(n); if (b < 0) throw std::runtime_error("broad phase body not found: " + n); g.bpBodies.push_back(b); }
  g.computeForwardKinematics();
  return g;
}

Graph Graph::fromDesc(const aff_scene_desc& d, const std::vector<std::string>& bodyNames,
                      const std::vector<std::string>& jointNames, const std::string& graphName)
{
  if ((int)bodyNames.size() != d.n_bodies || (int)jointNames.size() != d.n_joints) throw std::runtime_error("fromDesc: name lists do not match the description");
  Graph g;
  g.name = graphName;
  auto get12 = [](const double* p) { HTr A; for (int k = 0; k < 3; ++k) A.org[k] = p[k]; for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) A.rot[i][k] = p[3 + 3 * i + k]; return A; };
  g.bodies.resize((size_t)d.n_bodies); g.joints.resize((size_t)d.n_joints); g.shapes.resize((size_t)d.n_shapes);
  for (int b = 0; b < d.n_bodies; ++b) {
    Body& B = g.bodies[(size_t)b];
    B.name = bodyNames[(size_t)b]; B.parent = d.body_parent[b]; B.rigid_body_joints = d.body_rbj[b] != 0;
    B.A_BP = get12(d.body_A_BP + 12 * (size_t)b); B.A_BI = HTr::identity();
    B.firstJoint = d.body_first_joint[b]; B.nJoints = d.body_n_joints[b]; B.firstShape = d.body_first_shape[b]; B.nShapes = d.body_n_shapes[b];
    g.bodyIndex_[B.name] = b;
    for (int k = 0; k < B.nJoints; ++k) g.joints.at((size_t)(B.firstJoint + k)).body = b;
    for (int k = 0; k < B.nShapes; ++k) g.shapes.at((size_t)(B.firstShape + k)).body = b;
  }
  for (int j = 0; j < d.n_joints; ++j) {
    Joint& J = g.joints[(size_t)j];
    J.name = jointNames[(size_t)j]; J.type = d.joint_type[j]; J.constrained = d.joint_constrained[j] != 0;
    J.q_init = d.joint_q_init[j]; J.q0 = d.joint_q0[j]; J.q_min = d.joint_q_min[j]; J.q_max = d.joint_q_max[j];
    J.speedLimit = d.joint_speed_limit[j]; J.accLimit = d.joint_acc_limit[j]; J.weightJL = d.joint_weight_jl[j]; J.weightMetric = d.joint_weight_metric[j];
    J.A_JP = get12(d.joint_A_JP + 12 * (size_t)j);
    J.jacobiIndex = J.constrained ? -1 : g.nJ++;
    g.jointIndex_[J.name] = j;
  }
  for (int s = 0; s < d.n_shapes; ++s) {
    Shape& S = g.shapes[(size_t)s];
    S.type = d.shape_type[s]; S.distance = d.shape_distance[s] != 0; S.A_CB = get12(d.shape_A_CB + 12 * (size_t)s);
    for (int k = 0; k < 3; ++k) S.extents[k] = d.shape_extents[3 * (size_t)s + k];
  }
  g.bpThreshold = d.bp_threshold;
  g.bpTrees.assign(d.bp_tree_root, d.bp_tree_root + d.n_bp_trees);
  g.bpBodies.assign(d.bp_body, d.bp_body + d.n_bp_bodies);
  g.computeForwardKinematics();
  return g;
}

int Graph::getBodyByName(const std::string& n) const
{
  auto it = bodyIndex_.find(n);
  return it == bodyIndex_.end() ? -1 : it->second;
}

int Graph::getBodyByNameNoCase(const std::string& n) const
{
  int b = getBodyByName(n);
  if (b >= 0) return b;
  auto lower = [](std::string s) { for (char& c : s) c = (char)std::tolower((unsigned char)c); return s; };
  const std::string ln = lower(n);
  for (size_t i = 0; i < bodies.size(); ++i) if (lower(bodies[i].name) == ln) return (int)i;
  return -1;
}

int Graph::getJointByName(const std::string& n) const
{
  auto it = jointIndex_.find(n);
  return it == jointIndex_.end() ? -1 : it->second;
}

bool Graph::isChild(int body, int ancestor) const
{
  if (body < 0 || ancestor < 0) return false;
  for (int b = bodies[body].parent; b >= 0; b = bodies[b].parent) if (b == ancestor) return true;
  return false;
}

int Graph::numDistanceShapes(int body) const
{
  int n = 0;
  const Body& b = bodies[body];
  for (int k = 0; k < b.nShapes; ++k) if (shapes[b.firstShape + k].distance) ++n;
  return n;
}

std::vector<int> Graph::children(int body) const
{
  std::vector<int> c;
  for (size_t i = 0; i < bodies.size(); ++i) if (bodies[i].parent == body) c.push_back((int)i);
  return c;
}

static void applyJoint(HTr& A, int type, double q)
{
  if (type < AFF_JNT_ROT_X) { for (int k = 0; k < 3; ++k) A.org[k] += q * A.rot[type][k]; return; }
  const int d = type - AFF_JNT_ROT_X, a = (d + 1) % 3, b = (d + 2) % 3;
  const double s = std::sin(q), c = std::cos(q);
  for (int k = 0; k < 3; ++k) {
    const double ra = A.rot[a][k], rb = A.rot[b][k];
    A.rot[a][k] = c * ra + s * rb;
    A.rot[b][k] = -s * ra + c * rb;
  }
}

void Graph::computeForwardKinematics()
{
  // bodies may have been re-parented behind their index: iterate until settled
  std::vector<char> done(bodies.size(), 0);
  size_t nDone = 0;
  while (nDone < bodies.size()) {
    bool progressed = false;
    for (size_t i = 0; i < bodies.size(); ++i) {
      Body& b = bodies[i];
      if (done[i] || (b.parent >= 0 && !done[b.parent])) continue;
      HTr cur = b.parent >= 0 ? bodies[b.parent].A_BI : HTr::identity();
      for (int k = 0; k < b.nJoints; ++k) {
        const Joint& j = joints[b.firstJoint + k];
        cur = compose(j.A_JP, cur);
        applyJoint(cur, j.type, j.q_init);
      }
      b.A_BI = compose(b.A_BP, cur);
      done[i] = 1; ++nDone; progressed = true;
    }
    if (!progressed) throw std::runtime_error("kinematic graph has a cycle");
  }
}

void Graph::connectBody(int child, int parent, bool identityTransform)
{
  Body& c = bodies.at(child);
  if (!c.rigid_body_joints || c.nJoints < 6) throw std::runtime_error("connectBody: " + c.name + " has no rigid body joints");
  if (identityTransform) {      // setConnectTransform(HTr_identity()): the child sits on the parent frame
    for (int k = 0; k < 6; ++k) joints[c.firstJoint + k].q_init = 0.0;
    c.parent = parent;
    computeForwardKinematics();
    return;
  }
  const HTr& C = c.A_BI; const HTr& P = bodies.at(parent).A_BI;
  // relative pose -> six rigid-body dof values (x y z + x-y-z Euler angles)
  double rel[3], R[3][3];
  for (int i = 0; i < 3; ++i) {
    rel[i] = 0.0;
    for (int k = 0; k < 3; ++k) rel[i] += P.rot[i][k] * (C.org[k] - P.org[k]);
    for (int j = 0; j < 3; ++j) { R[i][j] = 0.0; for (int k = 0; k < 3; ++k) R[i][j] += C.rot[i][k] * P.rot[j][k]; }
  }
  double ea[3];
  const double cy = std::sqrt(R[0][0] * R[0][0] + R[1][0] * R[1][0]);
  if (cy > 1.0e-10) { ea[0] = std::atan2(-R[2][1], R[2][2]); ea[1] = std::atan2(R[2][0], cy); ea[2] = std::atan2(-R[1][0], R[0][0]); }
  else { ea[0] = std::atan2(R[1][2], R[1][1]); ea[1] = std::atan2(R[2][0], cy); ea[2] = 0.0; }
  for (int k = 0; k < 3; ++k) { joints[c.firstJoint + k].q_init = rel[k]; joints[c.firstJoint + 3 + k].q_init = ea[k]; }
  c.parent = parent;
  computeForwardKinematics();
}

void Graph::applyRollout(const aff_constraint* cons, int nCons, const double* qlog, int rows, double dt)
{
  if (!qlog || rows < 1) throw std::runtime_error("applyRollout: no joint log");
  auto setRow = [&](int row) {
    for (Joint& j : joints) if (j.jacobiIndex >= 0) j.q_init = qlog[(size_t)row * nJ + j.jacobiIndex];
    computeForwardKinematics();
  };
  // the time the trajectory controller has reached in iteration `iter` is dt added up iter + 1 times;
  // an event of iteration `iter` sees the poses of qlog row `iter` (TrajectoryPredictor.cpp:182-188)
  std::vector<char> fired((size_t)std::max(nCons, 0), 0);
  double tnow = 0.0;
  for (int iter = 0; iter + 1 < rows; ++iter) {
    tnow = tnow + dt;
    bool any = false;
    for (int k = 0; k < nCons; ++k) {
      if (fired[k]) continue;
      const aff_constraint& c = cons[k];
      if (c.kind == AFF_CON_CONNECT && c.t - tnow < 1.0e-8) any = true;
      if (c.kind == AFF_CON_COLLISION && c.t - tnow < 0.0) any = true;
    }
    if (!any) continue;
    setRow(iter);
    for (int k = 0; k < nCons; ++k) {
      if (fired[k]) continue;
      const aff_constraint& c = cons[k];
      if (c.kind == AFF_CON_CONNECT && c.t - tnow < 1.0e-8) { connectBody(c.body_a, c.body_b, c.dim == 1); fired[k] = 1; }
      else if (c.kind == AFF_CON_COLLISION && c.t - tnow < 0.0) {
        Body& b = bodies.at(c.body_a);
        for (int sI = 0; sI < b.nShapes; ++sI) shapes[b.firstShape + sI].distance = c.flag != 0;   // CollisionModelConstraint.h:101-147
        fired[k] = 1;
      }
    }
  }
  setRow(rows - 1);
  if (restoreParentFirstOrder()) computeForwardKinematics();
}

void Graph::rolloutTransforms(const aff_constraint* cons, int nCons, const double* qlog, int rows, double dt, double* out) const
{
  if (!qlog || rows < 1 || !out) throw std::runtime_error("rolloutTransforms: no joint log");
  Graph g = *this;                      // body ids stay those of the batch: no re-sort in here
  const size_t nb = g.bodies.size();
  auto setRow = [&](int row) {
    for (Joint& j : g.joints) if (j.jacobiIndex >= 0) j.q_init = qlog[(size_t)row * nJ + j.jacobiIndex];
    g.computeForwardKinematics();
  };
  auto record = [&](int row) {
    double* dst = out + (size_t)row * nb * 12;
    for (size_t b = 0; b < nb; ++b) g.bodies[b].A_BI.to12(dst + 12 * b);
  };
  setRow(0);
  record(0);
  std::vector<char> fired((size_t)std::max(nCons, 0), 0);
  double tnow = 0.0;
  for (int iter = 0; iter + 1 < rows; ++iter) {
    tnow = tnow + dt;
    // events of iteration `iter` act on the poses of row `iter` (still set), row iter + 1 sees their effect
    for (int k = 0; k < nCons; ++k) {
      if (fired[k]) continue;
      const aff_constraint& c = cons[k];
      if (c.kind == AFF_CON_CONNECT && c.t - tnow < 1.0e-8) { g.connectBody(c.body_a, c.body_b, c.dim == 1); fired[k] = 1; }
      else if (c.kind == AFF_CON_COLLISION && c.t - tnow < 0.0) fired[k] = 1;      // shape flags do not move anything
    }
    setRow(iter + 1);
    record(iter + 1);
  }
}

bool Graph::restoreParentFirstOrder()
{
  const int nb = (int)bodies.size();
  bool ok = true;
  for (int i = 0; i < nb && ok; ++i) ok = bodies[i].parent < i;
  if (ok) return false;
  // emit a body once its parent is out; bodies that are ready keep their relative order
  std::vector<int> order, newId(nb, -1);
  order.reserve(nb);
  std::vector<char> done(nb, 0);
  while ((int)order.size() < nb) {
    const size_t before = order.size();
    for (int i = 0; i < nb; ++i) {
      if (done[i]) continue;
      const int p = bodies[i].parent;
      if (p >= 0 && !done[p]) continue;
      done[i] = 1; newId[i] = (int)order.size(); order.push_back(i);
    }
    if (order.size() == before) throw std::runtime_error("graph has a parent cycle");
  }
  std::vector<Body> nbod;
  nbod.reserve(nb);
  for (int i : order) { Body b = bodies[i]; if (b.parent >= 0) b.parent = newId[b.parent]; nbod.push_back(b); }
  bodies.swap(nbod);
  for (Joint& j : joints) if (j.body >= 0) j.body = newId[j.body];
  for (Shape& sh : shapes) if (sh.body >= 0) sh.body = newId[sh.body];
  for (int& t : bpTrees) t = newId[t];
  for (int& t : bpBodies) t = newId[t];
  bodyIndex_.clear();
  for (int i = 0; i < nb; ++i) bodyIndex_[bodies[i].name] = i;
  return true;
}

static void shapeAABB(const Shape& s, const HTr& A, double mn[3], double mx[3])
{
  if (s.type == AFF_SHAPE_SSL || s.type == AFF_SHAPE_SPHERE) {
    for (int k = 0; k < 3; ++k) {
      const double p = A.org[k], q = p + (s.type == AFF_SHAPE_SSL ? s.extents[2] * A.rot[2][k] : 0.0);
      mn[k] = std::min(p, q) - s.extents[0]; mx[k] = std::max(p, q) + s.extents[0];
    }
  } else if (s.type == AFF_SHAPE_BOX) {
    for (int k = 0; k < 3; ++k) {
      const double e = 0.5 * (std::fabs(A.rot[0][k]) * s.extents[0] + std::fabs(A.rot[1][k]) * s.extents[1] + std::fabs(A.rot[2][k]) * s.extents[2]);
      mn[k] = A.org[k] - e; mx[k] = A.org[k] + e;
    }
  } else {
    for (int k = 0; k < 3; ++k) {
      const double az = A.rot[2][k];
      const double e = std::fabs(az) * 0.5 * s.extents[2] + s.extents[0] * std::sqrt(std::max(0.0, 1.0 - az * az));
      mn[k] = A.org[k] - e; mx[k] = A.org[k] + e;
    }
  }
}

bool Graph::computeBodyAABB(int body, double xyzMin[3], double xyzMax[3]) const
{
  const Body& b = bodies.at(body);
  bool any = false;
  for (int k = 0; k < b.nShapes; ++k) {
    const Shape& s = shapes[b.firstShape + k];
    if (!s.distance) continue;
    double mn[3], mx[3];
    shapeAABB(s, compose(s.A_CB, b.A_BI), mn, mx);
    for (int c = 0; c < 3; ++c) {
      xyzMin[c] = any ? std::min(xyzMin[c], mn[c]) : mn[c];
      xyzMax[c] = any ? std::max(xyzMax[c], mx[c]) : mx[c];
    }
    any = true;
  }
  return any;
}

// Ray against the world AABB of each distance shape: a conservative stand-in for
// Rcs' exact ray/shape tests, adequate for the "is anything above the object"
// query of ActionGet::init (reference src/ActionGet.cpp:244-261).
int Graph::closestRigidBodyInDirection(const double pt[3], const double dir[3], double* dMin) const
{
  int best = -1;
  double bestD = DBL_MAX;
  for (size_t i = 0; i < bodies.size(); ++i) {
    const Body& b = bodies[i];
    for (int k = 0; k < b.nShapes; ++k) {
      const Shape& s = shapes[b.firstShape + k];
      if (!s.distance) continue;
      double mn[3], mx[3];
      shapeAABB(s, compose(s.A_CB, b.A_BI), mn, mx);
      double t0 = 0.0, t1 = DBL_MAX;
      bool hit = true;
      for (int c = 0; c < 3 && hit; ++c) {
        if (std::fabs(dir[c]) < 1.0e-12) { if (pt[c] < mn[c] || pt[c] > mx[c]) hit = false; continue; }
        double a = (mn[c] - pt[c]) / dir[c], bb = (mx[c] - pt[c]) / dir[c];
        if (a > bb) std::swap(a, bb);
        t0 = std::max(t0, a); t1 = std::min(t1, bb);
        if (t0 > t1) hit = false;
      }
      if (hit && t0 < bestD) { bestD = t0; best = (int)i; }
    }
  }
  if (dMin) *dMin = bestD;
  return best;
}

const aff_scene_desc* Graph::toDesc()
{
  const size_t nb = bodies.size(), nj = joints.size(), ns = shapes.size();
  d_parent_.resize(nb); d_fj_.resize(nb); d_nj_.resize(nb); d_fs_.resize(nb); d_ns_.resize(nb); d_rbj_.resize(nb); d_abp_.resize(12 * nb);
  for (size_t i = 0; i < nb; ++i) {
    const Body& b = bodies[i];
    d_parent_[i] = b.parent; d_fj_[i] = b.firstJoint; d_nj_[i] = b.nJoints; d_fs_[i] = b.firstShape; d_ns_[i] = b.nShapes;
    d_rbj_[i] = b.rigid_body_joints ? 1 : 0;
    b.A_BP.to12(&d_abp_[12 * i]);
  }
  d_jtype_.resize(nj); d_jcon_.resize(nj); d_ajp_.resize(12 * nj); d_qinit_.resize(nj); d_q0_.resize(nj); d_qmin_.resize(nj);
  d_qmax_.resize(nj); d_speed_.resize(nj); d_acc_.resize(nj); d_wjl_.resize(nj); d_wm_.resize(nj);
  for (size_t i = 0; i < nj; ++i) {
    const Joint& j = joints[i];
    d_jtype_[i] = j.type; d_jcon_[i] = j.constrained ? 1 : 0; j.A_JP.to12(&d_ajp_[12 * i]);
    d_qinit_[i] = j.q_init; d_q0_[i] = j.q0; d_qmin_[i] = j.q_min; d_qmax_[i] = j.q_max;
    d_speed_[i] = j.speedLimit; d_acc_[i] = j.accLimit; d_wjl_[i] = j.weightJL; d_wm_[i] = j.weightMetric;
  }
  d_stype_.resize(ns); d_sdist_.resize(ns); d_acb_.resize(12 * ns); d_ext_.resize(3 * ns);
  for (size_t i = 0; i < ns; ++i) {
    const Shape& s = shapes[i];
    d_stype_[i] = s.type; d_sdist_[i] = s.distance ? 1 : 0; s.A_CB.to12(&d_acb_[12 * i]);
    for (int k = 0; k < 3; ++k) d_ext_[3 * i + k] = s.extents[k];
  }
  d_tree_.assign(bpTrees.begin(), bpTrees.end());
  d_bpb_.assign(bpBodies.begin(), bpBodies.end());
  aff_scene_desc& d = desc_;
  d.n_bodies = (int32_t)nb; d.n_joints = (int32_t)nj; d.n_shapes = (int32_t)ns;
  d.body_parent = d_parent_.data(); d.body_first_joint = d_fj_.data(); d.body_n_joints = d_nj_.data();
  d.body_first_shape = d_fs_.data(); d.body_n_shapes = d_ns_.data(); d.body_rbj = d_rbj_.data(); d.body_A_BP = d_abp_.data();
  d.joint_type = d_jtype_.data(); d.joint_constrained = d_jcon_.data(); d.joint_A_JP = d_ajp_.data();
  d.joint_q_init = d_qinit_.data(); d.joint_q0 = d_q0_.data(); d.joint_q_min = d_qmin_.data(); d.joint_q_max = d_qmax_.data();
  d.joint_speed_limit = d_speed_.data(); d.joint_acc_limit = d_acc_.data(); d.joint_weight_jl = d_wjl_.data();
  d.joint_weight_metric = d_wm_.data();
  d.shape_type = d_stype_.data(); d.shape_distance = d_sdist_.data(); d.shape_A_CB = d_acb_.data(); d.shape_extents = d_ext_.data();
  d.bp_threshold = bpThreshold;
  d.n_bp_trees = (int32_t)d_tree_.size(); d.bp_tree_root = d_tree_.data();
  d.n_bp_bodies = (int32_t)d_bpb_.size(); d.bp_body = d_bpb_.data();
  // bodies the elbow / wrist null-space terms look up (reference src/TrajectoryPredictor.cpp:459-463,492-493,540-542)
  d.ns_base = getBodyByName("base_footprint");
  d.ns_upperarm[0] = getBodyByName("upperarm_left"); d.ns_upperarm[1] = getBodyByName("upperarm_right");
  d.ns_forearm[0] = getBodyByName("forearm_left"); d.ns_forearm[1] = getBodyByName("forearm_right");
  d.ns_hand[0] = getBodyByName("hand_left"); d.ns_hand[1] = getBodyByName("hand_right");
  return &desc_;
}

}   // namespace affgpu

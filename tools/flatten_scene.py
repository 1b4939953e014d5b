#!/usr/bin/env python3
"""Flatten an AffAction scene XML (Rcs graph + affordance layer) into the
line-based `.affscene` text format read by affaction_b200/csrc/host/Graph.cpp.

The reference parses these files twice: Rcs' graph parser builds the
kinematic tree (bodies, joints, shapes, model state, broad phase), and
`ActionScene::parse` (reference src/ActionScene.cpp:191-258) reads the
<AffordanceModel>/<Manipulator> layer.  Rcs itself is not vendored in the
reference, so the graph semantics below follow SURVEY.md Appendix B; each
rule is spelled out next to the code that implements it.

This is a build-time tool: /root/reference does not exist on the GPU box, so
its outputs are committed under affaction_b200/data/.

Usage: flatten_scene.py <scene.xml> <out.affscene> [search_dir ...]
"""
import math
import os
import sys
import xml.etree.ElementTree as ET

XI_INCLUDE = "{http://www.w3.org/2001/XInclude}include"
XI_INCLUDE_2003 = "{http://www.w3.org/2003/XInclude}include"

JOINT_TYPES = {"TransX": 0, "TransY": 1, "TransZ": 2, "RotX": 3, "RotY": 4, "RotZ": 5}
SHAPE_TYPES = {"SSL": 0, "SPHERE": 1, "BOX": 2, "CYLINDER": 3}
BIG = 1.0e308  # stands in for DBL_MAX (unlimited speed / acceleration)


def euler_to_rot(a, b, c):
    """x-y-z Euler angles (rad) -> A_KI (rows = axes of K in I coordinates)."""
    sa, ca = math.sin(a), math.cos(a)
    sb, cb = math.sin(b), math.cos(b)
    sc, cc = math.sin(c), math.cos(c)
    return [
        [cb * cc, ca * sc + sa * sb * cc, sa * sc - ca * sb * cc],
        [-cb * sc, ca * cc - sa * sb * sc, sa * cc + ca * sb * sc],
        [sb, -sa * cb, ca * cb],
    ]


def identity():
    return [0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]


def parse_transform(txt):
    """'x y z a b c' (metres, degrees) -> 12 doubles: org(3) + rot(9 row-major)."""
    if txt is None:
        return identity()
    v = [float(t) for t in txt.split()]
    if len(v) != 6:
        raise ValueError("transform needs 6 values: %r" % txt)
    r = euler_to_rot(math.radians(v[3]), math.radians(v[4]), math.radians(v[5]))
    return [v[0], v[1], v[2]] + r[0] + r[1] + r[2]


def compose(t_child, t_parent):
    """A_2I = A_21 o A_1I for 12-double transforms (child given relative to parent)."""
    po, pr = t_parent[:3], [t_parent[3:6], t_parent[6:9], t_parent[9:12]]
    co, cr = t_child[:3], [t_child[3:6], t_child[6:9], t_child[9:12]]
    org = [po[i] + sum(pr[k][i] * co[k] for k in range(3)) for i in range(3)]
    rot = [[sum(cr[i][k] * pr[k][j] for k in range(3)) for j in range(3)] for i in range(3)]
    return org + rot[0] + rot[1] + rot[2]


def parse_bool(txt, default):
    if txt is None:
        return default
    return txt.strip().lower() in ("true", "1", "yes")


class Flattener:
    def __init__(self, search_dirs):
        self.search_dirs = search_dirs
        self.bodies = []          # dicts in document order
        self.body_index = {}
        self.model_states = []    # (model, [(joint, deg)])
        self.model_state_stamps = []
        self.broadphase = None
        self.entities = []
        self.manipulators = []
        self.graph_name = None

    # ---------------------------------------------------------------- XML walk
    def load(self, path):
        tree = ET.parse(path)
        root = tree.getroot()
        self.graph_name = root.get("name", "")
        self.walk(root, os.path.dirname(os.path.abspath(path)), "", None, None, top=True)

    def resolve(self, href, base_dir):
        cands = [os.path.join(base_dir, href)] + [os.path.join(d, href) for d in self.search_dirs]
        cands += [os.path.join(d, os.path.basename(href)) for d in self.search_dirs]
        for c in cands:
            if os.path.isfile(c):
                return c
        raise FileNotFoundError("cannot resolve include %r from %s" % (href, base_dir))

    def walk(self, node, base_dir, suffix, group_prev, group_tf, top=False):
        """Visit the children of a <Graph> or <Group> in document order.

        Group rules: the group's `name` is appended to
        every body / joint name and `prev` reference inside it; a body without
        `prev` attaches to the group's `prev`, with the group's `transform`
        pre-multiplied to the body's own.
        """
        for ch in node:
            tag = ch.tag
            if tag in (XI_INCLUDE, XI_INCLUDE_2003):
                inc = self.resolve(ch.get("href"), base_dir)
                sub = ET.parse(inc).getroot()
                self.walk(sub, os.path.dirname(inc), suffix, group_prev, group_tf)
            elif tag == "Graph":
                self.walk(ch, base_dir, suffix, group_prev, group_tf)
            elif tag == "Group":
                g_suffix = ch.get("name", "") + suffix
                g_prev = ch.get("prev")
                g_prev = (g_prev + suffix) if g_prev is not None else group_prev
                # a group's prev refers to a body outside of it: try suffixed, else plain
                if ch.get("prev") is not None and g_prev not in self.body_index:
                    g_prev = ch.get("prev")
                g_tf = parse_transform(ch.get("transform")) if ch.get("transform") else None
                if g_tf is not None and group_tf is not None and ch.get("prev") is None:
                    g_tf = compose(g_tf, group_tf)
                elif g_tf is None and ch.get("prev") is None:
                    g_tf = group_tf
                self.walk(ch, base_dir, g_suffix, g_prev, g_tf)
            elif tag == "Body":
                self.add_body(ch, suffix, group_prev, group_tf)
            elif tag == "model_state":
                if top:
                    js = [(j.get("joint"), float(j.get("position"))) for j in ch if j.tag == "joint_state"]
                    self.model_states.append((ch.get("model"), js))
                    self.model_state_stamps.append(int(ch.get("time_stamp", "0")))
            elif tag == "BroadPhase":
                if top:
                    self.broadphase = {
                        "threshold": float(ch.get("DistanceThreshold", "0")),
                        "trees": [t.get("root") for t in ch if t.tag == "Tree"],
                        "bodies": [t.get("name") for t in ch if t.tag == "Body"],
                    }
            elif tag == "AffordanceModel":
                self.add_entity(ch)
            elif tag == "Manipulator":
                self.add_manipulator(ch)
            # ActionSequence, Agent, Component, noShape ... : not on the predict path

    def add_body(self, node, suffix, group_prev, group_tf):
        name = node.get("name") + suffix
        prev = node.get("prev")
        tf = parse_transform(node.get("transform"))
        if prev is not None:
            pname = prev + suffix
            if pname not in self.body_index:
                pname = prev  # reference to a body outside the group
            if pname not in self.body_index:
                raise KeyError("body %s: unknown prev %s" % (name, prev))
            parent = self.body_index[pname]
        elif group_prev is not None:
            parent = self.body_index[group_prev]
            if group_tf is not None:
                tf = compose(tf, group_tf)
        else:
            parent = -1
        joints = []
        rbj_attr = node.get("rigid_body_joints")
        has_rbj = False
        if rbj_attr is not None and rbj_attr.strip().lower() not in ("false", "0"):
            has_rbj = True
            vals = rbj_attr.split()
            if len(vals) == 6:
                q6 = [float(v) for v in vals]
            else:
                # rigid_body_joints="true": the body transform (if any) moves into the six dofs
                q6 = [0.0] * 6
                if node.get("transform"):
                    q6 = [float(v) for v in node.get("transform").split()]
                    tf = identity()
            for k, jt in enumerate(("TransX", "TransY", "TransZ", "RotX", "RotY", "RotZ")):
                q = q6[k] if k < 3 else math.radians(q6[k])
                joints.append(dict(name="%s_rigidBodyJnt%d" % (name, k), type=JOINT_TYPES[jt], constrained=1,
                                   q_init=q, q0=q, q_min=-BIG, q_max=BIG, speed=BIG, acc=BIG,
                                   wjl=1.0, wmetric=1.0, A_JP=identity()))
        shapes = []
        for ch in node:
            if ch.tag == "Joint":
                joints.append(self.parse_joint(ch, suffix))
            elif ch.tag == "Shape":
                sh = self.parse_shape(ch)
                if sh is not None:
                    shapes.append(sh)
        body = dict(name=name, parent=parent, rbj=int(has_rbj), A_BP=tf, joints=joints, shapes=shapes)
        self.body_index[name] = len(self.bodies)
        self.bodies.append(body)

    def parse_joint(self, node, suffix):
        jt = node.get("type")
        rot = jt.startswith("Rot")
        conv = math.radians if rot else (lambda x: x)
        rng = [float(v) for v in node.get("range", "0").split()]
        if len(rng) == 1:
            q_min, q0, q_max = -abs(rng[0]), 0.0, abs(rng[0])
        elif len(rng) == 2:
            q_min, q_max = rng
            q0 = 0.5 * (q_min + q_max)
        else:
            q_min, q0, q_max = rng[:3]
        speed = node.get("speedLimit")
        acc = node.get("accelerationLimit")
        return dict(
            name=node.get("name") + suffix, type=JOINT_TYPES[jt],
            constrained=int(parse_bool(node.get("constraint"), False)),
            q_init=conv(q0), q0=conv(q0), q_min=conv(q_min), q_max=conv(q_max),
            speed=conv(float(speed)) if speed is not None else BIG,
            acc=conv(float(acc)) if acc is not None else BIG,
            wjl=float(node.get("weightJL", "1")), wmetric=float(node.get("weightMetric", "1")),
            A_JP=parse_transform(node.get("transform")))

    def parse_shape(self, node):
        """Only primitive shapes that can ever carry RCSSHAPE_COMPUTE_DISTANCE are kept.

        MESH and FRAME shapes are skipped: all meshes in the shipped scenes are
        distance="false", and CollisionModelConstraint never enables them
        (reference src/CollisionModelConstraint.h:122-127).  `distance`
        defaults to true when the attribute is absent.
        """
        st = node.get("type")
        if st not in SHAPE_TYPES:
            return None
        dist = parse_bool(node.get("distance"), True)
        ext = [0.0, 0.0, 0.0]
        if st == "BOX":
            ext = [float(v) for v in node.get("extents", "0 0 0").split()]
        else:
            r = float(node.get("radius", "0"))
            ln = float(node.get("length", "0"))
            ext = [r, r, ln]
        return dict(type=SHAPE_TYPES[st], dist=int(dist), A_CB=parse_transform(node.get("transform")), ext=ext)

    def add_entity(self, node):
        body = node.get("body")
        ent = dict(name=node.get("name", body), body=body, type=node.get("type", "-"), affs=[])
        for ch in node:
            if ch.tag.startswith("xxx") or ch.tag == "Ingredient":
                continue
            kv = {k: v for k, v in ch.attrib.items() if k != "frame"}
            ent["affs"].append(dict(cls=ch.tag, frame=ch.get("frame", body), kv=kv))
        self.entities.append(ent)

    def add_manipulator(self, node):
        m = dict(name=node.get("name"), type=node.get("type", "-"), fingers=[], caps=[])
        for ch in node:
            if ch.tag == "Joints":
                m["fingers"] = ch.get("names", "").split()
            elif ch.tag.endswith("Capability"):
                m["caps"].append(dict(cls=ch.tag, frame=ch.get("frame", "").replace(" ", ",")))
        self.manipulators.append(m)

    # ---------------------------------------------------------------- emit
    def apply_model_state(self):
        jmap = {}
        for b in self.bodies:
            for j in b["joints"]:
                jmap[j["name"]] = j
        for model, js in self.model_states:
            if model != self.graph_name:
                continue
            for jn, deg in js:
                j = jmap.get(jn)
                if j is None:
                    continue
                # The model state becomes the graph's default state: both the start
                # pose and the centre q0 the joint-limit cost pulls towards
                # (RcsGraph_changeDefaultState; see DESIGN.md "restatement choices").
                j["q_init"] = math.radians(deg) if j["type"] >= 3 else deg
                j["q0"] = j["q_init"]

    def emit(self, out):
        self.apply_model_state()
        f = lambda v: " ".join(repr(float(x)) for x in v)
        w = out.write
        n_joints = sum(len(b["joints"]) for b in self.bodies)
        n_shapes = sum(len(b["shapes"]) for b in self.bodies)
        w("AFFSCENE 1\n")
        w("graph %s\n" % (self.graph_name or "-"))
        w("counts %d %d %d\n" % (len(self.bodies), n_joints, n_shapes))
        for i, b in enumerate(self.bodies):
            w("body %d %s %d %d %d %d %s\n" % (i, b["name"], b["parent"], b["rbj"], len(b["joints"]),
                                              len(b["shapes"]), f(b["A_BP"])))
            for j in b["joints"]:
                w("joint %s %d %d %s %s\n" % (j["name"], j["type"], j["constrained"],
                                              f([j["q_init"], j["q0"], j["q_min"], j["q_max"], j["speed"], j["acc"],
                                                 j["wjl"], j["wmetric"]]), f(j["A_JP"])))
            for s in b["shapes"]:
                w("shape %d %d %s %s\n" % (s["type"], s["dist"], f(s["ext"]), f(s["A_CB"])))
        bp = self.broadphase or dict(threshold=0.0, trees=[], bodies=[])
        w("broadphase %s %d %d\n" % (repr(bp["threshold"]), len(bp["trees"]), len(bp["bodies"])))
        for t in bp["trees"]:
            w("tree %s\n" % t)
        for t in bp["bodies"]:
            w("bpbody %s\n" % t)
        w("entities %d\n" % len(self.entities))
        for e in self.entities:
            w("entity %s %s %s %d\n" % (e["name"], e["body"], e["type"], len(e["affs"])))
            for a in e["affs"]:
                kv = " ".join("%s=%s" % (k, v.replace(" ", ",")) for k, v in sorted(a["kv"].items()))
                w("aff %s %s %d %s\n" % (a["cls"], a["frame"], len(a["kv"]), kv))
        w("manipulators %d\n" % len(self.manipulators))
        for m in self.manipulators:
            w("manipulator %s %s %d %d\n" % (m["name"], m["type"], len(m["fingers"]), len(m["caps"])))
            for fj in m["fingers"]:
                w("finger %s\n" % fj)
            for c in m["caps"]:
                w("cap %s %s\n" % (c["cls"], c["frame"]))
        # every <model_state> by name and time stamp (RcsGraph_getModelStateFromXML: the "pose" action)
        jmap = {j["name"]: j for b in self.bodies for j in b["joints"]}
        w("modelstates %d\n" % len(self.model_states))
        for (model, js), stamp in zip(self.model_states, self.model_state_stamps):
            known = [(jn, v) for jn, v in js if jn in jmap]
            w("modelstate %s %d %d\n" % (model, stamp, len(known)))
            for jn, v in known:
                w("msjoint %s %s\n" % (jn, repr(math.radians(v) if jmap[jn]["type"] >= 3 else v)))
        w("end\n")


def main(argv):
    if len(argv) < 3:
        print(__doc__)
        return 2
    fl = Flattener(argv[3:])
    fl.load(argv[1])
    with open(argv[2], "w") as out:
        fl.emit(out)
    nj = sum(len(b["joints"]) for b in fl.bodies)
    print("%s: %d bodies, %d joints (dof), %d unconstrained, %d kept shapes (%d distance)" % (
        argv[2], len(fl.bodies), nj,
        sum(1 for b in fl.bodies for j in b["joints"] if not j["constrained"]),
        sum(len(b["shapes"]) for b in fl.bodies),
        sum(1 for b in fl.bodies for s in b["shapes"] if s["dist"])))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))

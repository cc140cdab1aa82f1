"""twin.py — a SECOND, independent CPU restatement of the SbFDDP node model (numpy).  TEST INFRASTRUCTURE ONLY: nothing
under eagle-mpc_b200/ imports it; only tests/ and scripts/make_twin_golden.py do.

Why it exists.  oracle/oracle.cpp and the CUDA kernels were written from the same recalled formulas and consume the same
flattened problem (`empc_problem_desc_t`) produced by the product's host mirror; an error shared by both would be
invisible to every GPU-vs-oracle test.  This twin shares NOTHING with them:
  * front-end: the YAML files are read with PyYAML and the URDFs with xml.etree, straight from disk — not through
    host/params.cpp, host/urdf.cpp or host/trajectory.cpp — and turned into stages / cost tables / the knot layout by
    following the reference's own files (src/trajectory.cpp:102-143, src/stage.cpp:52-70, src/factory/cost.cpp,
    src/factory/activation.cpp:20-100, src/multicopter-base-params.cpp:67-101, src/sbfddp.cpp:169-190,464-477);
  * dynamics: forward dynamics as M(q)^-1 (tau - h(q, v)) with M and h from the recursive Newton-Euler algorithm in body
    coordinates (the oracle and the kernels run the articulated-body algorithm and a world-frame derivative recursion);
  * derivatives: every Jacobian (Fx, Fu, Lx, Lu and the residual Jacobians inside the Gauss-Newton Hessians) by the
    COMPLEX STEP (h = 1e-30) of the twin's own calc, as SURVEY.md Appendix B.8 recommends — no analytic derivative code
    at all, so nothing to get wrong twice;
  * solver algebra: one dense Riccati sweep and one rollout written directly from SURVEY.md section 8 (a5)-(a8).
Every function works on complex arrays (no abs / atan2 / norm; branches look at real parts only).
"""
import os
import xml.etree.ElementTree as ET

import numpy as np
import yaml as pyyaml

H = 1e-30  # complex step

# ---------------------------------------------------------------------------------------------------------------------
# Lie groups.  Tangent ordering [linear; angular].


def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.result_type(w, float))


def _abc(t2):
    """A = sin t / t, B = (1 - cos t) / t^2, C = (t - sin t) / t^3 as functions of t^2 (series near zero)"""
    if abs(t2) < 1e-6:
        return (1 - t2 / 6 + t2 * t2 / 120, 0.5 - t2 / 24 + t2 * t2 / 720, 1.0 / 6 - t2 / 120 + t2 * t2 / 5040)
    t = np.sqrt(t2)
    return np.sin(t) / t, (1 - np.cos(t)) / t2, (t - np.sin(t)) / (t2 * t)


def exp3(w):
    t2 = w @ w
    A, B, _ = _abc(t2)
    W = hat(w)
    return np.eye(3) + A * W + B * (W @ W)


def log3(R):
    c = (R[0, 0] + R[1, 1] + R[2, 2] - 1) / 2
    s = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2   # sin(t) * axis
    s2 = s @ s
    if c.real > 0.5:
        if abs(s2) < 1e-6:  # asin(x)/x = 1 + x^2/6 + 3 x^4/40 + 5 x^6/112 + 35 x^8/1152
            f = 1 + s2 / 6 + 3 * s2**2 / 40 + 5 * s2**3 / 112 + 35 * s2**4 / 1152
        else:
            sn = np.sqrt(s2)
            f = np.arcsin(sn) / sn
        return f * s
    if c.real > -0.9:
        sn = np.sqrt(s2)
        return (np.arccos(c) / sn) * s
    # near pi the skew part vanishes: take the axis from the symmetric part, R_sym = I + (1 - cos t)(n n^T - I),
    # its sign from the skew part (s = sin t n), and t = pi - asin(n . s)
    N = np.eye(3) + ((R + R.T) / 2 - np.eye(3)) / (1 - c)
    i = int(np.argmax([N[0, 0].real, N[1, 1].real, N[2, 2].real]))
    n = N[:, i] / np.sqrt(N[i, i])
    if (n @ s).real < 0:
        n = -n
    return (np.pi - np.arcsin(n @ s)) * n


def exp6(nu):
    v, w = nu[:3], nu[3:]
    t2 = w @ w
    A, B, C = _abc(t2)
    W = hat(w)
    W2 = W @ W
    return np.eye(3) + A * W + B * W2, (np.eye(3) + B * W + C * W2) @ v


def log6(R, p):
    w = log3(R)
    t2 = w @ w
    _, B, C = _abc(t2)
    W = hat(w)
    V = np.eye(3) + B * W + C * (W @ W)
    return np.concatenate([np.linalg.solve(V, p), w])


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def quat_of_rotvec(w):
    t2 = w @ w
    if abs(t2) < 1e-6:
        sh = 0.5 - t2 / 48 + t2 * t2 / 3840          # sin(t/2)/t
        ch = 1 - t2 / 8 + t2 * t2 / 384              # cos(t/2)
    else:
        t = np.sqrt(t2)
        sh, ch = np.sin(t / 2) / t, np.cos(t / 2)
    return np.concatenate([sh * w, [ch]])


def rpy_to_R(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr], [-sp, cp * sr, cp * cr]])


# ---------------------------------------------------------------------------------------------------------------------
# Robot model from a URDF, the way pinocchio::urdf::buildModel(path, JointModelFreeFlyer()) sees it


class Robot:
    def __init__(self, urdf_path):
        root = ET.parse(urdf_path).getroot()
        links = {l.get("name"): l for l in root.findall("link")}
        joints = root.findall("joint")
        children = {j.find("child").get("link") for j in joints}
        base = [n for n in links if n not in children]
        assert len(base) == 1, "one root link expected"
        # joint 0: free-flyer carrying the root link
        self.parent, self.place_R, self.place_p, self.axis = [-1], [np.eye(3)], [np.zeros(3)], [np.zeros(3)]
        self.mass, self.mc, self.I_o = [0.0], [np.zeros(3)], [np.zeros((3, 3))]   # body inertia about the JOINT origin
        self.effort = [0.0]
        self.frames = {}   # link name -> (joint index, R, p) placement in the joint frame

        def xyz_rpy(el):
            o = el.find("origin") if el is not None else None
            if o is None:
                return np.eye(3), np.zeros(3)
            xyz = np.array([float(v) for v in (o.get("xyz") or "0 0 0").split()])
            rpy = [float(v) for v in (o.get("rpy") or "0 0 0").split()]
            return rpy_to_R(*rpy), xyz

        def add_body(j, link, R, p):
            """add the inertial of `link`, whose frame sits at (R, p) in joint j's frame, to joint j's body"""
            self.frames[link.get("name")] = (j, R, p)
            inertial = link.find("inertial")
            if inertial is None:
                return
            m = float(inertial.find("mass").get("value"))
            Rc, pc = xyz_rpy(inertial)
            i = inertial.find("inertia")
            Ic = np.array([[float(i.get("ixx")), float(i.get("ixy")), float(i.get("ixz"))],
                           [float(i.get("ixy")), float(i.get("iyy")), float(i.get("iyz"))],
                           [float(i.get("ixz")), float(i.get("iyz")), float(i.get("izz"))]])
            c = p + R @ pc                       # centre of mass in the joint frame
            Ic_j = (R @ Rc) @ Ic @ (R @ Rc).T    # rotational inertia about the COM, joint axes
            C = hat(c)
            self.mass[j] += m
            self.mc[j] = self.mc[j] + m * c
            self.I_o[j] = self.I_o[j] + Ic_j - m * (C @ C)   # parallel axis: about the joint origin

        def walk(link_name, j, R, p):
            add_body(j, links[link_name], R, p)
            for jn in joints:
                if jn.find("parent").get("link") != link_name:
                    continue
                Rj, pj = xyz_rpy(jn)
                child = jn.find("child").get("link")
                if jn.get("type") == "fixed":
                    walk(child, j, R @ Rj, p + R @ pj)
                elif jn.get("type") in ("revolute", "continuous"):
                    ax = np.array([float(v) for v in jn.find("axis").get("xyz").split()])
                    self.parent.append(j); self.place_R.append(R @ Rj); self.place_p.append(p + R @ pj)
                    self.axis.append(ax / np.sqrt(ax @ ax))
                    self.mass.append(0.0); self.mc.append(np.zeros(3)); self.I_o.append(np.zeros((3, 3)))
                    lim = jn.find("limit")
                    self.effort.append(float(lim.get("effort")) if lim is not None and lim.get("effort") else 0.0)
                    walk(child, len(self.parent) - 1, np.eye(3), np.zeros(3))
                else:
                    raise ValueError("joint type " + jn.get("type"))

        walk(base[0], 0, np.eye(3), np.zeros(3))
        self.nj = len(self.parent)
        self.na = self.nj - 1
        self.nq, self.nv = 7 + self.na, 6 + self.na
        self.nx, self.ndx = self.nq + self.nv, 2 * self.nv
        self.gravity = np.array([0.0, 0.0, -9.81])

    # ---- kinematics: local joint transforms and world placements ----
    def joint_transforms(self, q):
        """(R_i, p_i) of joint i's frame in its parent's frame (joint 0: in the world)"""
        out = [(quat_to_R(q[3:7]), q[0:3])]
        for i in range(1, self.nj):
            out.append((self.place_R[i] @ exp3(self.axis[i] * q[6 + i]), self.place_p[i] + 0 * q[0]))
        return out

    def world_placements(self, q):
        T = self.joint_transforms(q)
        oM = [T[0]]
        for i in range(1, self.nj):
            Rp, pp = oM[self.parent[i]]
            oM.append((Rp @ T[i][0], pp + Rp @ T[i][1]))
        return oM

    def _inertia_apply(self, i, mot):
        v, w = mot[:3], mot[3:]
        f = self.mass[i] * v - np.cross(self.mc[i], w)
        n = np.cross(self.mc[i], v) + self.I_o[i] @ w
        return np.concatenate([f, n])

    def rnea(self, q, v, a, gravity=True):
        """inverse dynamics in body coordinates; motions [lin; ang], forces [force; torque]"""
        T = self.joint_transforms(q)
        ct = np.result_type(q, v, a, float)
        vel, acc, frc = [None] * self.nj, [None] * self.nj, [None] * self.nj
        g = self.gravity if gravity else np.zeros(3)
        for i in range(self.nj):
            R, p = T[i]
            if i == 0:
                vp = np.zeros(6, dtype=ct)
                ap = np.concatenate([-g, np.zeros(3)]).astype(ct)   # a_0 = -gravity (world)
                S_qd, S_qdd = v[0:6], a[0:6]
            else:
                vp, ap = vel[self.parent[i]], acc[self.parent[i]]
                S = np.concatenate([np.zeros(3), self.axis[i]])
                S_qd, S_qdd = S * v[5 + i], S * a[5 + i]
            # X^-1 on motions: (R^T (v - p x w), R^T w)
            vi = np.concatenate([R.T @ (vp[:3] - np.cross(p, vp[3:])), R.T @ vp[3:]])
            ai = np.concatenate([R.T @ (ap[:3] - np.cross(p, ap[3:])), R.T @ ap[3:]])
            vel[i] = vi + S_qd
            # motion cross product v_i x (S qd): (w x v' + v x w', w x w')
            vx = vel[i]
            cross = np.concatenate([np.cross(vx[3:], S_qd[:3]) + np.cross(vx[:3], S_qd[3:]), np.cross(vx[3:], S_qd[3:])])
            acc[i] = ai + S_qdd + cross
            hI = self._inertia_apply(i, vel[i])
            fI = self._inertia_apply(i, acc[i])
            # v x* h: (w x f, w x n + v x f)
            frc[i] = fI + np.concatenate([np.cross(vx[3:], hI[:3]), np.cross(vx[3:], hI[3:]) + np.cross(vx[:3], hI[:3])])
        tau = np.zeros(self.nv, dtype=ct)
        for i in range(self.nj - 1, -1, -1):
            if i == 0:
                tau[0:6] = frc[0]
            else:
                tau[5 + i] = self.axis[i] @ frc[i][3:]
                R, p = T[i]
                f, n = R @ frc[i][:3], R @ frc[i][3:]
                par = self.parent[i]
                frc[par] = frc[par] + np.concatenate([f, n + np.cross(p, f)])
        return tau, vel

    def mass_and_bias(self, q, v):
        z = np.zeros(self.nv)
        h, vel = self.rnea(q, v, z)
        M = np.zeros((self.nv, self.nv), dtype=np.result_type(q, float))
        for k in range(self.nv):
            e = np.zeros(self.nv); e[k] = 1.0
            M[:, k] = self.rnea(q, z, e, gravity=False)[0]
        return M, h, vel

    def forward_dynamics(self, q, v, tau):
        M, h, vel = self.mass_and_bias(q, v)
        return np.linalg.solve(M, tau - h), vel

    def frame_motion(self, frame, q, v, a):
        """LOCAL spatial velocity and (gravity-free) spatial acceleration of an operational frame, [lin; ang], for joint
        velocities v and accelerations a: the kinematic half of the body-coordinate recursion, then jMf.actInv"""
        T = self.joint_transforms(q)
        ct = np.result_type(q, v, a, float)
        vel, acc = [None] * self.nj, [None] * self.nj
        for i in range(self.nj):
            R, p = T[i]
            if i == 0:
                vel[0] = v[0:6].astype(ct); acc[0] = a[0:6].astype(ct)   # v x v = 0
                continue
            vp, ap = vel[self.parent[i]], acc[self.parent[i]]
            S = np.concatenate([np.zeros(3), self.axis[i]])
            vi = np.concatenate([R.T @ (vp[:3] - np.cross(p, vp[3:])), R.T @ vp[3:]])
            ai = np.concatenate([R.T @ (ap[:3] - np.cross(p, ap[3:])), R.T @ ap[3:]])
            vel[i] = vi + S * v[5 + i]
            vx, sq = vel[i], S * v[5 + i]
            acc[i] = ai + S * a[5 + i] + np.concatenate([np.cross(vx[3:], sq[:3]) + np.cross(vx[:3], sq[3:]), np.cross(vx[3:], sq[3:])])
        j, Rf, pf = self.frames[frame]
        to_f = lambda m: np.concatenate([Rf.T @ (m[:3] - np.cross(pf, m[3:])), Rf.T @ m[3:]])
        return to_f(vel[j]), to_f(acc[j])

    def contact_acceleration(self, contact, q, v, a):
        """what the contact constrains to zero (crocoddyl ContactModel3D / 6D, zero gains): the classical linear
        acceleration of the frame origin in LOCAL coordinates (3D), the LOCAL spatial acceleration (6D)"""
        vf, af = self.frame_motion(contact["frame"], q, v, a)
        if contact["type"] == "ContactModel3D":
            return af[:3] + np.cross(vf[3:], vf[:3])
        return af

    def contact_dynamics(self, contact, q, v, tau):
        """pinocchio::forwardDynamics(q, v, tau, Jc, a0, 0): [M Jc^T; Jc 0] [a; -lambda] = [tau - h; -a0]"""
        M, h, vel = self.mass_and_bias(q, v)
        z = np.zeros(self.nv)
        a0 = self.contact_acceleration(contact, q, v, z)
        nc = a0.size
        Jc = np.zeros((nc, self.nv), dtype=M.dtype)
        for k in range(self.nv):   # the constrained acceleration is affine in a
            e = np.zeros(self.nv); e[k] = 1.0
            Jc[:, k] = self.contact_acceleration(contact, q, v, e) - a0
        K = np.block([[M, Jc.T], [Jc, np.zeros((nc, nc))]])
        sol = np.linalg.solve(K, np.concatenate([tau - h, -a0]))
        return sol[:self.nv], -sol[self.nv:], vel


# ---------------------------------------------------------------------------------------------------------------------
# State manifold


def integrate(rob, x, dx):
    q, v = x[:rob.nq], x[rob.nq:]
    R = quat_to_R(q[3:7])
    Re, t = exp6(dx[0:6])
    quat = quat_mul(q[3:7], quat_of_rotvec(dx[3:6]))
    quat = quat / np.sqrt(quat @ quat)
    qn = np.concatenate([q[0:3] + R @ t, quat, q[7:] + dx[6:rob.nv]])
    return np.concatenate([qn, v + dx[rob.nv:]])


def diff(rob, x0, x1):
    q0, q1 = x0[:rob.nq], x1[:rob.nq]
    R0, R1 = quat_to_R(q0[3:7]), quat_to_R(q1[3:7])
    base = log6(R0.T @ R1, R0.T @ (q1[0:3] - q0[0:3]))
    return np.concatenate([base, q1[7:] - q0[7:], x1[rob.nq:] - x0[rob.nq:]])


# ---------------------------------------------------------------------------------------------------------------------
# Problem front-end: YAML -> platform, stages, cost tables, knot layout

COST_DIM = {"CostModelState": None, "CostModelControl": None, "CostModelFramePlacement": 6, "CostModelFrameRotation": 3,
            "CostModelFrameVelocity": 6, "CostModelFrameTranslation": 3, "CostModelContactFrictionCone": 5}
BIG = np.finfo(float).max   # crocoddyl::FrictionCone's "infinite" bound


def _vec(v):
    return np.array([float(a) for a in v], dtype=float)


class Problem:
    """trajectory YAML -> what SolverSbFDDP sees (robot, actuation, one cost dict per stage, node -> stage map)"""

    def __init__(self, yaml_path, yaml_root, urdf_root, dt_ms, barrier_weight=1e-3, integrator="IntegratedActionModelEuler",
                 use_squash=True):
        # use_squash = False: the problem the reference hands to crocoddyl::SolverBoxFDDP (plain multicopter actuation, no
        # barrier cost; src/trajectory.cpp:96-100, examples/python/trajectory.py:19-24)
        self.use_squash = use_squash
        assert integrator in ("IntegratedActionModelEuler", "IntegratedActionModelRK4")   # src/factory/int-action.cpp:24-35
        self.rk4 = integrator == "IntegratedActionModelRK4"
        doc = pyyaml.safe_load(open(os.path.join(yaml_root, yaml_path)))["trajectory"]
        self.rob = Robot(os.path.join(urdf_root, doc["robot"]["urdf"]))
        plat = pyyaml.safe_load(open(os.path.join(yaml_root, doc["robot"]["follow"])))["platform"]
        rob = self.rob
        # src/multicopter-base-params.cpp:67-101
        cf, cm = float(plat["cf"]), float(plat["cm"])
        rotors = plat["$rotors"]
        self.nr = int(plat["n_rotors"])
        assert len(rotors) == self.nr
        self.tau_f = np.zeros((6, self.nr))
        for i, r in enumerate(rotors):
            qr = _vec(r["orientation"]); qr = qr / np.linalg.norm(qr)
            th = quat_to_R(qr) @ np.array([0.0, 0.0, 1.0])
            self.tau_f[0:3, i] = th
            self.tau_f[3:6, i] = np.cross(_vec(r["translation"]), th) + float(r["spin_direction"][0]) * cm / cf * th
        self.nu = self.nr + rob.na
        self.u_lb = np.concatenate([np.full(self.nr, float(plat["min_thrust"])), -np.array(rob.effort[1:])])
        self.u_ub = np.concatenate([np.full(self.nr, float(plat["max_thrust"])), np.array(rob.effort[1:])])
        self.dt = dt_ms / 1000.0
        self.barrier_weight = barrier_weight
        x0 = doc.get("initial_state")
        self.x0 = _vec(x0) if x0 is not None else np.concatenate([[0, 0, 0, 0, 0, 0, 1], np.zeros(rob.na + rob.nv)])
        # stages (src/stage.cpp:52-70; costs live in a std::map: iterated by name)
        self.stages, node_stage = [], []
        last0 = False
        for si, st in enumerate(doc["stages"]):
            costs = {}
            for c in st["costs"]:
                costs[c["name"]] = self._cost(c)
            # src/stage.cpp:38-47, src/factory/contacts.cpp:32-81 (one contact per stage in the corpus)
            contacts = st.get("contacts") or []
            assert len(contacts) <= 1, "the twin handles one contact per stage"
            contact = None
            for c in contacts:
                assert c["link_name"] in self.rob.frames, "Link " + c["link_name"] + " does not exist"
                gains = _vec(c["gains"]) if "gains" in c else np.zeros(2)
                assert not gains.any(), "Baumgarte gains: not restated in the twin"
                contact = {"type": c["type"], "frame": c["link_name"]}
            # knot rule, src/trajectory.cpp:117-127 (integer division)
            dur = int(st["duration"])
            if dur // dt_ms == 0 and si + 1 < len(doc["stages"]):
                n_knots, last0 = 1, True
            else:
                n_knots = dur // dt_ms - (1 if last0 else 0)
                last0 = False
            node_stage += [si] * n_knots
            # SolverSbFDDP::barrierInit (src/sbfddp.cpp:181-186): every RUNNING model gets "barrier" (one model object per
            # stage, src/trajectory.cpp:134: a last stage without running knots is the terminal model only and has none)
            if n_knots > 0 and use_squash:
                costs["barrier"] = {"type": "Barrier", "weight": barrier_weight, "active": True}
            self.stages.append({"name": st["name"], "duration": dur, "costs": dict(sorted(costs.items())), "contact": contact})
        self.T = len(node_stage)
        self.node_stage = node_stage + [len(self.stages) - 1]   # terminal model = the last stage's model (:135)

    def _cost(self, c):
        rob = self.rob
        t = c["type"]
        out = {"type": t, "weight": float(c["weight"]), "active": "active" not in c}   # presence of the key => inactive
        if t == "CostModelState":
            nr = rob.ndx
            out["ref"] = _vec(c["reference"]) if "reference" in c else np.concatenate([[0, 0, 0, 0, 0, 0, 1], np.zeros(rob.na + rob.nv)])
        elif t == "CostModelControl":
            nr = self.nu
            out["ref"] = _vec(c["reference"]) if "reference" in c else np.zeros(self.nu)
        elif t == "CostModelContactFrictionCone":
            # crocoddyl::FrictionCone(n_surf, mu, nf = 4, inner_appr = false) (src/factory/cost.cpp:149-167): rows
            # (-mu z +- t_i)^T c_R_o for the nf/2 tangents t_i, then n_surf^T; bounds (-inf, 0] and [0, +inf)
            nr = 5
            n = _vec(c["n_surf"]); n = n / np.linalg.norm(n)
            mu = float(c["mu"])
            z = np.array([0.0, 0.0, 1.0])
            ax = np.cross(n, z)   # rotation taking n_surf onto z (Quaternion::FromTwoVectors)
            sn, cs = np.linalg.norm(ax), n @ z
            cRo = np.eye(3) if sn < 1e-12 else exp3(ax / sn * np.arctan2(sn, cs))
            A = np.zeros((5, 3))
            for i in range(2):
                th = 2 * np.pi * i / 4
                ts = np.array([np.cos(th), np.sin(th), 0.0])
                A[2 * i] = (-mu * z + ts) @ cRo
                A[2 * i + 1] = (-mu * z - ts) @ cRo
            A[4] = n
            out["A"] = A
            out["frame"] = c["link_name"]
            out["act"] = "ActivationModelQuadraticBarrier"
            out["lb"] = np.array([-BIG] * 4 + [0.0]); out["ub"] = np.array([0.0] * 4 + [BIG])
            return out
        else:
            nr = COST_DIM[t]
            out["frame"] = c["link_name"]
            assert c["link_name"] in rob.frames, "Link " + c["link_name"] + " does not exist"
            if t in ("CostModelFramePlacement", "CostModelFrameRotation"):
                qo = _vec(c["orientation"]); qo = qo / np.linalg.norm(qo)
                out["R"] = quat_to_R(qo)
            if t in ("CostModelFramePlacement", "CostModelFrameTranslation"):
                out["p"] = _vec(c["position"])
            if t == "CostModelFrameVelocity":
                out["vref"] = np.concatenate([_vec(c["linear"]), _vec(c["angular"])])
        act = c.get("activation", "ActivationModelQuad")   # src/factory/activation.cpp:25-30
        out["act"] = act
        if act in ("ActivationModelWeightedQuad", "ActivationModelWeightedQuadraticBarrier"):
            out["w"] = _vec(c["weights"]) if "weights" in c else np.ones(nr)
            assert out["w"].size == nr
        if act in ("ActivationModelQuadraticBarrier", "ActivationModelWeightedQuadraticBarrier"):
            out["lb"], out["ub"] = _vec(c["l_bound"]), _vec(c["u_bound"])
            assert out["lb"].size == nr and out["ub"].size == nr
        return out

    # ---- node model ----
    def squash(self, u, smooth):
        if not self.use_squash:
            return u
        a = (smooth * (self.u_ub - self.u_lb)) ** 2
        return 0.5 * (np.sqrt((u - self.u_lb) ** 2 + a) - np.sqrt((u - self.u_ub) ** 2 + a) + self.u_lb + self.u_ub)

    def dynamics(self, stage, x, u, smooth):
        """(joint accelerations, contact force or None): DifferentialActionModelFreeFwdDynamics, or
        DifferentialActionModelContactFwdDynamics for a stage with a contact (src/factory/diff-action.cpp:30-35)"""
        rob = self.rob
        q, v = x[:rob.nq], x[rob.nq:]
        s = self.squash(u, smooth)
        tau = np.concatenate([self.tau_f @ s[:self.nr], s[self.nr:]])
        contact = self.stages[stage]["contact"]
        if contact is None:
            return rob.forward_dynamics(q, v, tau)[0], None
        a, lam, _ = rob.contact_dynamics(contact, q, v, tau)
        return a, lam

    def residuals(self, stage, x, u, smooth):
        """[(name, weight, r, activation dict)] of the active costs of the stage, in the reference's iteration order"""
        rob = self.rob
        q, v = x[:rob.nq], x[rob.nq:]
        oM = None
        out = []
        lam = None
        for name, c in self.stages[stage]["costs"].items():
            if not c["active"]:
                continue
            t = c["type"]
            if t == "Barrier":   # WeightedQuadraticBarrier on the pre-squash control (src/sbfddp.cpp:171-176,464-477)
                w = 1.0 / (smooth * (self.u_ub - self.u_lb)) ** 2
                out.append((name, c["weight"], u, {"act": "ActivationModelWeightedQuadraticBarrier", "w": w,
                                                            "lb": self.u_lb, "ub": self.u_ub}))
                continue
            if t == "CostModelState":
                r = diff(rob, c["ref"].astype(x.dtype), x)
            elif t == "CostModelControl":
                r = u - c["ref"]
            elif t == "CostModelContactFrictionCone":   # r = A f, f the contact force in the contact frame
                if lam is None:
                    lam = self.dynamics(stage, x, u, smooth)[1]
                assert lam is not None and self.stages[stage]["contact"]["frame"] == c["frame"]
                r = c["A"] @ lam[:3]
            else:
                if oM is None:
                    oM = rob.world_placements(q)
                    vel = rob.rnea(q, v, np.zeros(rob.nv))[1]
                j, Rf, pf = rob.frames[c["frame"]]
                oR, op = oM[j][0] @ Rf, oM[j][1] + oM[j][0] @ pf
                if t == "CostModelFramePlacement":
                    r = log6(c["R"].T @ oR, c["R"].T @ (op - c["p"]))
                elif t == "CostModelFrameRotation":
                    r = log3(c["R"].T @ oR)
                elif t == "CostModelFrameTranslation":
                    r = op - c["p"]
                else:  # LOCAL frame velocity: jMf.actInv(v_j)
                    vj = vel[j]
                    r = np.concatenate([Rf.T @ (vj[:3] - np.cross(pf, vj[3:])), Rf.T @ vj[3:]]) - c["vref"]
            out.append((name, c["weight"], r, c))
        return out

    @staticmethod
    def activation(c, r):
        """(value, Arr diagonal)"""
        act = c["act"]
        if act == "ActivationModelQuad":
            return 0.5 * (r @ r), np.ones(r.size)
        if act == "ActivationModelWeightedQuad":
            return 0.5 * (r @ (c["w"] * r)), c["w"]
        lo, hi = r - c["lb"], r - c["ub"]
        rl = np.where(lo.real < 0, lo, 0)
        ru = np.where(hi.real > 0, hi, 0)
        on = ((lo.real <= 0) | (hi.real >= 0)).astype(float)
        if act == "ActivationModelQuadraticBarrier":
            return 0.5 * (rl @ rl) + 0.5 * (ru @ ru), on
        w = c["w"]   # upstream: value and gradient carry w^2, the Gauss-Newton weight is w
        return 0.5 * ((w * rl) @ (w * rl)) + 0.5 * ((w * ru) @ (w * ru)), on * w

    def node_terms(self, stage, x, u, smooth):
        """(xnext, [(total weight, residual, activation dict)]) of one node.  IntegratedActionModelEuler: semi-implicit
        Euler, costs at (x, u) weighted by dt.  IntegratedActionModelRK4 (crocoddyl/core/integrator/rk4.hxx): stage states
        y_i = x (+) c_i dt k_{i-1}, k_i = [v(y_i); a(y_i, u)], c = (0, 1/2, 1/2, 1), dx = dt/6 (k_0 + 2 k_1 + 2 k_2 + k_3), and
        the costs of the four stages weighted by dt/6 (1, 2, 2, 1)."""
        rob = self.rob
        if not self.rk4:
            a, _lam = self.dynamics(stage, x, u, smooth)
            dx = np.concatenate([x[rob.nq:] * self.dt + a * self.dt ** 2, a * self.dt])
            return integrate(rob, x, dx), [(self.dt * w, r, c) for _n, w, r, c in self.residuals(stage, x, u, smooth)]
        terms, ks = [], []
        y = x
        for i, (ci, wi) in enumerate(((0.0, 1.0), (0.5, 2.0), (0.5, 2.0), (1.0, 1.0))):
            if i:
                y = integrate(rob, x, ci * self.dt * ks[-1])
            a, _lam = self.dynamics(stage, y, u, smooth)
            ks.append(np.concatenate([y[rob.nq:], a]))
            terms += [(self.dt / 6.0 * wi * w, r, c) for _n, w, r, c in self.residuals(stage, y, u, smooth)]
        dx = self.dt / 6.0 * (ks[0] + 2 * ks[1] + 2 * ks[2] + ks[3])
        return integrate(rob, x, dx), terms

    def calc(self, stage, x, u, smooth, terminal=False):
        """(xnext, cost) of the integrated action model's calc; the terminal node evaluates with u = 0 (1.x convention)"""
        if terminal:
            u = np.zeros(self.nu, dtype=x.dtype)
        xnext, terms = self.node_terms(stage, x, u, smooth)
        cost = 0
        for wt, r, c in terms:
            cost = cost + wt * self.activation(c, r)[0]
        return xnext, cost

    def calc_diff(self, stage, x, u, smooth, terminal=False):
        """dict of the node's blocks by the complex step: Fx, Fu, Lx, Lu and the Gauss-Newton Lxx, Luu, Lxu (sum over the
        node's residual terms of w R^T Arr R with the TOTAL residual Jacobians: for RK4 that is exactly rk4.hxx's
        dyi_dx^T Lxx_i dyi_dx / ... assembly), plus xnext, cost"""
        rob = self.rob
        ndx, nu = rob.ndx, self.nu
        x = x.astype(complex); u = (np.zeros(nu) if terminal else u).astype(complex)
        xn0, c0 = self.calc(stage, x, u, smooth, terminal)
        res0 = self.node_terms(stage, x, u, smooth)[1]
        Fx, Fu = np.zeros((ndx, ndx)), np.zeros((ndx, nu))
        Lx, Lu = np.zeros(ndx), np.zeros(nu)
        Rx = [np.zeros((r.size, ndx)) for _w, r, _c in res0]
        Ru = [np.zeros((r.size, nu)) for _w, r, _c in res0]
        for k in range(ndx):
            e = np.zeros(ndx, dtype=complex); e[k] = 1j * H
            xk = integrate(rob, x, e)
            xn, c = self.calc(stage, xk, u, smooth, terminal)
            Fx[:, k] = diff(rob, xn0.real.astype(complex), xn).imag / H
            Lx[k] = c.imag / H
            for i, (_w, r, _c) in enumerate(self.node_terms(stage, xk, u, smooth)[1]):
                Rx[i][:, k] = r.imag / H
        if not terminal:
            for k in range(nu):
                uk = u.copy(); uk[k] += 1j * H
                xn, c = self.calc(stage, x, uk, smooth)
                Fu[:, k] = diff(rob, xn0.real.astype(complex), xn).imag / H
                Lu[k] = c.imag / H
                for i, (_w, r, _c) in enumerate(self.node_terms(stage, x, uk, smooth)[1]):
                    Ru[i][:, k] = r.imag / H
        Lxx, Luu, Lxu = np.zeros((ndx, ndx)), np.zeros((nu, nu)), np.zeros((ndx, nu))
        for i, (wt, r, c) in enumerate(res0):
            arr = self.activation(c, r)[1]
            Lxx += wt * Rx[i].T @ (arr[:, None] * Rx[i])
            Luu += wt * Ru[i].T @ (arr[:, None] * Ru[i])
            Lxu += wt * Rx[i].T @ (arr[:, None] * Ru[i])
        return {"xnext": xn0.real, "cost": c0.real, "Fx": Fx, "Fu": Fu, "Lx": Lx, "Lu": Lu, "Lxx": Lxx, "Luu": Luu, "Lxu": Lxu}


# ---------------------------------------------------------------------------------------------------------------------
# Solver algebra, dense


def box_qp(H, q, lb, ub, xinit, maxiter=100, th_acceptstep=0.1, th_grad=1e-5, reg=0.0):
    """Projected-Newton box QP (Tassa, Mansard, Todorov 2014, as crocoddyl::BoxQP runs it): min 1/2 x'Hx + q'x, lb <= x <= ub.
    Returns x, the free and clamped index arrays of the last active-set pass and the inverse of the free block of H."""
    x = np.clip(np.asarray(xinit, dtype=float), lb, ub)
    f = lambda v: 0.5 * v @ H @ v + q @ v
    free = np.arange(x.size); clamped = np.array([], dtype=int); Hff_inv = None
    for it in range(maxiter):
        g = q + H @ x
        on = ((x == lb) & (g > 0)) | ((x == ub) & (g < 0))
        clamped, free = np.flatnonzero(on), np.flatnonzero(~on)
        done = np.abs(g).max() <= th_grad or free.size == 0
        if not done or it == 0 or Hff_inv.shape[0] != free.size:   # (a stale inverse of another free set is never handed out)
            Hff = H[np.ix_(free, free)] + reg * np.eye(free.size)
            np.linalg.cholesky(Hff) if free.size else None   # raises when not positive definite (the reference's backward_error)
            Hff_inv = np.linalg.inv(Hff) if free.size else np.zeros((0, 0))
        if done:
            break
        dx = np.zeros_like(x)
        dx[free] = Hff_inv @ (-q[free] - H[np.ix_(free, clamped)] @ x[clamped]) - x[free]
        fold = f(x)
        for n in range(10):
            xnew = np.clip(x + dx / 2.0 ** n, lb, ub)
            if fold - f(xnew) > th_acceptstep * (g @ (x - xnew)):
                x = xnew
                break
    return x, free, clamped, Hff_inv


def riccati_sweep(nodes, terminal, fs, xreg, feasible, box=None):
    """nodes[t]: dict with Fx, Fu, Lx, Lu, Lxx, Lxu, Luu; terminal: Lx, Lxx; fs[t]: gaps (T + 1).  Returns K, k, Vx, Vxx.
    box = (us, u_lb, u_ub, k_prev): the gains of crocoddyl's SolverBoxFDDP / SolverBoxDDP for a feasible candidate — the box QP
    on du with u_lb <= us + du <= u_ub warm-started at k_prev, feedback from the free block only, clamped entries of Qu
    zeroed; then the sweep also returns the list of Qu."""
    T = len(nodes)
    ndx = terminal["Lx"].size
    Vxx = [None] * (T + 1); Vx = [None] * (T + 1); K = [None] * T; k = [None] * T; Qus = [None] * T
    Vxx[T] = terminal["Lxx"] + xreg * np.eye(ndx)
    Vx[T] = terminal["Lx"].copy()
    if not feasible:
        Vx[T] = Vx[T] + Vxx[T] @ fs[T]
    for t in range(T - 1, -1, -1):
        n = nodes[t]
        Fx, Fu = n["Fx"], n["Fu"]
        Qxx = n["Lxx"] + Fx.T @ Vxx[t + 1] @ Fx
        Qxu = n["Lxu"] + Fx.T @ Vxx[t + 1] @ Fu
        Quu = n["Luu"] + Fu.T @ Vxx[t + 1] @ Fu + xreg * np.eye(Fu.shape[1])
        Qx = n["Lx"] + Fx.T @ Vx[t + 1]
        Qu = n["Lu"] + Fu.T @ Vx[t + 1]
        if box is not None and feasible:
            us, u_lb, u_ub, k_prev = box
            du, free, clamped, Hff_inv = box_qp(Quu, Qu, u_lb - us[t], u_ub - us[t], k_prev[t])
            Quu_inv = np.zeros_like(Quu)
            Quu_inv[np.ix_(free, free)] = Hff_inv
            K[t] = Quu_inv @ Qxu.T
            k[t] = -du
            Qu = Qu.copy(); Qu[clamped] = 0.0
            Qus[t] = Qu
        else:
            K[t] = np.linalg.solve(Quu, Qxu.T)
            k[t] = np.linalg.solve(Quu, Qu)
        Vx[t] = Qx + K[t].T @ (Quu @ k[t]) - 2 * K[t].T @ Qu
        V = Qxx - Qxu @ K[t]
        Vxx[t] = 0.5 * (V + V.T) + xreg * np.eye(ndx)
        if not feasible:
            Vx[t] = Vx[t] + Vxx[t] @ fs[t]
    if box is not None:
        return K, k, Vx, Vxx, Qus
    return K, k, Vx, Vxx


def rollout(prob, x0, xs, us, K, k, fs, alpha, smooth, feasible, clamp=None):
    """SolverFDDP::forwardPass from x0 with step length alpha: returns xs_try, us_try, cost_try.
    clamp = (u_lb, u_ub): SolverBoxFDDP::forwardPass clamps every trial control to its limits."""
    rob = prob.rob
    T = len(us)
    xs_try, us_try, cost = [], [], 0.0
    xnext = np.asarray(x0, dtype=float)
    for t in range(T + 1):
        xt = xnext if (feasible or alpha == 1) else integrate(rob, xnext, fs[t] * (alpha - 1))
        xs_try.append(xt)
        if t == T:
            cost += prob.calc(prob.node_stage[T], xt, None, smooth, terminal=True)[1]
            break
        ut = us[t] - alpha * k[t] - K[t] @ diff(rob, xs[t], xt)
        if clamp is not None:
            ut = np.clip(ut, clamp[0], clamp[1])
        us_try.append(ut)
        xnext, c = prob.calc(prob.node_stage[t], xt, ut, smooth)
        cost += c
    return np.array(xs_try), np.array(us_try), cost

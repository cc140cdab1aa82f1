// urdf.cpp — minimal URDF reader producing what pinocchio::urdf::buildModel(path, JointModelFreeFlyer(), model) gives
// eagle-mpc (src/trajectory.cpp:29-31, src/mpc-base.cpp:24-26):
//   joint 1 = free-flyer "root_joint" carrying the root link's inertia, one 1-DoF joint per revolute/continuous URDF
//   joint (placement = <origin> relative to the parent joint frame, axis from <axis>), fixed joints merged into the
//   parent body, a BODY frame per link, effortLimit from <limit effort="..">.
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

#include "eagle_mpc.hpp"

namespace eagle_mpc {
namespace {

struct Xml {
  std::string name;
  std::map<std::string, std::string> attr;
  std::vector<Xml> children;
  const Xml* child(const std::string& n) const {
    for (auto& c : children) if (c.name == n) return &c;
    return nullptr;
  }
};

struct XmlParser {
  const std::string& s;
  size_t p = 0;
  explicit XmlParser(const std::string& str) : s(str) {}
  void skip_ws() { while (p < s.size() && std::isspace((unsigned char)s[p])) ++p; }
  bool starts(const char* t) const { return s.compare(p, std::strlen(t), t) == 0; }
  void skip_misc() {
    for (;;) {
      skip_ws();
      if (starts("<?")) { p = s.find("?>", p); p = (p == std::string::npos) ? s.size() : p + 2; }
      else if (starts("<!--")) { p = s.find("-->", p); p = (p == std::string::npos) ? s.size() : p + 3; }
      else if (starts("<!")) { p = s.find(">", p); p = (p == std::string::npos) ? s.size() : p + 1; }
      else break;
    }
  }
  bool parse_element(Xml& out) {
    skip_misc();
    if (p >= s.size() || s[p] != '<' || starts("</")) return false;
    ++p;
    size_t b = p;
    while (p < s.size() && !std::isspace((unsigned char)s[p]) && s[p] != '>' && s[p] != '/') ++p;
    out.name = s.substr(b, p - b);
    for (;;) {
      skip_ws();
      if (p >= s.size()) throw std::runtime_error("URDF: unexpected end of file");
      if (s[p] == '/') { p += 2; return true; }
      if (s[p] == '>') { ++p; break; }
      b = p;
      while (p < s.size() && s[p] != '=' && !std::isspace((unsigned char)s[p])) ++p;
      const std::string key = s.substr(b, p - b);
      skip_ws(); ++p; skip_ws();
      const char q = s[p++];
      b = p;
      while (p < s.size() && s[p] != q) ++p;
      out.attr[key] = s.substr(b, p - b);
      ++p;
    }
    for (;;) {
      skip_misc();
      if (p >= s.size()) throw std::runtime_error("URDF: unterminated element " + out.name);
      if (starts("</")) { p = s.find('>', p) + 1; return true; }
      if (s[p] == '<') { Xml c; if (parse_element(c)) out.children.push_back(c); }
      else { while (p < s.size() && s[p] != '<') ++p; }  // text content ignored
    }
  }
};

std::vector<double> nums(const std::string& s, size_t n, double def = 0) {
  std::vector<double> v;
  std::stringstream ss(s);
  double x;
  while (ss >> x) v.push_back(x);
  v.resize(n, def);
  return v;
}
struct Se3 { double R[9]; double p[3]; };
Se3 identity() { Se3 m{{1, 0, 0, 0, 1, 0, 0, 0, 1}, {0, 0, 0}}; return m; }
Se3 from_origin(const Xml* origin) {
  Se3 m = identity();
  if (!origin) return m;
  auto xyz = nums(origin->attr.count("xyz") ? origin->attr.at("xyz") : "", 3);
  auto rpy = nums(origin->attr.count("rpy") ? origin->attr.at("rpy") : "", 3);
  const double cr = std::cos(rpy[0]), sr = std::sin(rpy[0]), cp = std::cos(rpy[1]), sp = std::sin(rpy[1]),
               cy = std::cos(rpy[2]), sy = std::sin(rpy[2]);
  const double R[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                       sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                       -sp, cp * sr, cp * cr};  // Rz(yaw) Ry(pitch) Rx(roll)
  std::memcpy(m.R, R, sizeof(R));
  for (int i = 0; i < 3; ++i) m.p[i] = xyz[i];
  return m;
}
Se3 mul(const Se3& A, const Se3& B) {
  Se3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C.R[3 * i + j] = A.R[3 * i] * B.R[j] + A.R[3 * i + 1] * B.R[3 + j] + A.R[3 * i + 2] * B.R[6 + j];
  for (int i = 0; i < 3; ++i) C.p[i] = A.R[3 * i] * B.p[0] + A.R[3 * i + 1] * B.p[1] + A.R[3 * i + 2] * B.p[2] + A.p[i];
  return C;
}
struct Body { double m = 0; double mc[3] = {0, 0, 0}; double Io[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; };  // inertia about the joint origin
void add_body(Body& b, double m, const Se3& M /* inertial frame in joint frame */, const double* Ic /* 3x3 in inertial frame */) {
  // rotate Ic into the joint frame, shift to the joint origin (parallel axis)
  double RI[9], I[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) RI[3 * i + j] = M.R[3 * i] * Ic[j] + M.R[3 * i + 1] * Ic[3 + j] + M.R[3 * i + 2] * Ic[6 + j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) I[3 * i + j] = RI[3 * i] * M.R[3 * j] + RI[3 * i + 1] * M.R[3 * j + 1] + RI[3 * i + 2] * M.R[3 * j + 2];
  const double* c = M.p;
  const double c2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) b.Io[3 * i + j] += I[3 * i + j] + m * ((i == j ? c2 : 0.0) - c[i] * c[j]);
  b.m += m;
  for (int i = 0; i < 3; ++i) b.mc[i] += m * c[i];
}
}  // namespace

std::size_t RobotModel::getFrameId(const std::string& name) const {
  for (std::size_t i = 0; i < frames.size(); ++i) if (frames[i].name == name) return i;
  return frames.size();
}

VectorXd StateMultibody::zero() const {
  VectorXd x(get_nx(), 0.0);
  x[6] = 1.0;
  return x;
}

std::shared_ptr<RobotModel> buildModelFromUrdf(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::invalid_argument("The file " + path + " does not contain a valid URDF model.");
  std::stringstream ss; ss << f.rdbuf();
  const std::string text = ss.str();
  XmlParser xp(text);
  Xml robot;
  if (!xp.parse_element(robot) || robot.name != "robot") throw std::invalid_argument("The file " + path + " does not contain a valid URDF model.");

  std::map<std::string, const Xml*> links;
  std::map<std::string, std::vector<const Xml*>> child_joints;  // parent link -> joints
  std::map<std::string, bool> is_child;
  for (auto& c : robot.children) {
    if (c.name == "link") links[c.attr.at("name")] = &c;
    if (c.name == "joint") {
      child_joints[c.child("parent")->attr.at("link")].push_back(&c);
      is_child[c.child("child")->attr.at("link")] = true;
    }
  }
  std::string root;
  for (auto& c : robot.children)
    if (c.name == "link" && !is_child.count(c.attr.at("name"))) { root = c.attr.at("name"); break; }
  if (root.empty()) throw std::invalid_argument("URDF has no root link: " + path);

  auto model = std::make_shared<RobotModel>();
  std::vector<Body> bodies;
  auto add_link_inertia = [&](int joint, const Se3& link_in_joint, const Xml* link) {
    const Xml* in = link->child("inertial");
    if (!in) return;
    const double m = in->child("mass") ? std::stod(in->child("mass")->attr.at("value")) : 0.0;
    double Ic[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (const Xml* I = in->child("inertia")) {
      auto g = [&](const char* k) { return I->attr.count(k) ? std::stod(I->attr.at(k)) : 0.0; };
      Ic[0] = g("ixx"); Ic[1] = Ic[3] = g("ixy"); Ic[2] = Ic[6] = g("ixz"); Ic[4] = g("iyy"); Ic[5] = Ic[7] = g("iyz"); Ic[8] = g("izz");
    }
    add_body(bodies[joint], m, mul(link_in_joint, from_origin(in->child("origin"))), Ic);
  };
  // depth-first traversal in document order (pinocchio visits children in URDF order)
  struct Item { std::string link; int joint; Se3 link_in_joint; };
  std::vector<Item> stack;
  model->parent.push_back(-1);
  model->joint_names.push_back("root_joint");
  { Se3 I = identity(); model->jplace_R.push_back(std::vector<double>(I.R, I.R + 9)); model->jplace_p.push_back({0, 0, 0}); model->axis.push_back({0, 0, 1}); }
  bodies.emplace_back();
  model->effortLimit.assign(6, 0.0);
  stack.push_back({root, 0, identity()});
  std::vector<Item> order;
  while (!stack.empty()) {
    Item it = stack.back(); stack.pop_back();
    order.push_back(it);
    add_link_inertia(it.joint, it.link_in_joint, links.at(it.link));
    model->frames.push_back({it.link, it.joint, std::vector<double>(it.link_in_joint.R, it.link_in_joint.R + 9),
                             std::vector<double>(it.link_in_joint.p, it.link_in_joint.p + 3)});
    std::vector<Item> kids;
    for (const Xml* j : child_joints[it.link]) {
      const std::string type = j->attr.at("type"), child = j->child("child")->attr.at("link");
      const Se3 Mj = mul(it.link_in_joint, from_origin(j->child("origin")));
      if (type == "fixed") kids.push_back({child, it.joint, Mj});
      else if (type == "revolute" || type == "continuous") {
        const int idx = (int)model->parent.size();
        model->parent.push_back(it.joint);
        model->joint_names.push_back(j->attr.at("name"));
        model->jplace_R.push_back(std::vector<double>(Mj.R, Mj.R + 9));
        model->jplace_p.push_back(std::vector<double>(Mj.p, Mj.p + 3));
        auto ax = nums(j->child("axis") ? j->child("axis")->attr.at("xyz") : "1 0 0", 3);
        const double n = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
        model->axis.push_back({ax[0] / n, ax[1] / n, ax[2] / n});
        bodies.emplace_back();
        const Xml* lim = j->child("limit");
        model->effortLimit.push_back(lim && lim->attr.count("effort") ? std::stod(lim->attr.at("effort")) : 0.0);
        kids.push_back({child, idx, identity()});
      } else throw std::invalid_argument("URDF joint type '" + type + "' is not supported");
    }
    for (auto k = kids.rbegin(); k != kids.rend(); ++k) stack.push_back(*k);
  }
  model->njoints = (int)model->parent.size();
  model->nq = 7 + model->njoints - 1;
  model->nv = 6 + model->njoints - 1;
  for (const Body& b : bodies) {
    std::vector<double> c(3, 0.0), I(9, 0.0);
    if (b.m > 0) {
      for (int i = 0; i < 3; ++i) c[i] = b.mc[i] / b.m;
      const double c2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) I[3 * i + j] = b.Io[3 * i + j] - b.m * ((i == j ? c2 : 0.0) - c[i] * c[j]);
    }
    model->mass.push_back(b.m); model->com.push_back(c); model->inertia.push_back(I);
  }
  return model;
}

}  // namespace eagle_mpc

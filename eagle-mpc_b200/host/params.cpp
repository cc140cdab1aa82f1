// params.cpp — YAML-subset reader, ParserYaml key flattening, converter<T>, path helpers.
//
// Mirrors src/utils/parser_yaml.cpp (key layout, `follow:` includes, `$key` atomic maps, "transition" presence
// semantics :274-278), src/utils/converter_utils.cpp (vector literals: no scientific notation inside [...], trailing
// comma tolerated) and include/eagle_mpc/utils/params_server.hpp.  yaml-cpp is not available here, so a small
// indentation-based parser covers the YAML subset the eagle-mpc corpus uses: block maps, block sequences, flow
// sequences (possibly continued on following lines), quoted scalars and comments.
#include <dlfcn.h>
#include <limits.h>

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <regex>
#include <sstream>

#include "eagle_mpc.hpp"

namespace eagle_mpc {

// ---- paths ----------------------------------------------------------------------------------------------------------
static std::string g_yaml_dir, g_robot_dir;
static std::string repo_root() {
  // <root>/eagle-mpc_b200/lib/libempc_b200.so -> <root>
  Dl_info info;
  std::string f = ".";
  if (dladdr((void*)&set_yaml_dir, &info) && info.dli_fname) f = info.dli_fname;
  char buf[4096];
  if (realpath(f.c_str(), buf)) f = buf;
  for (int i = 0; i < 3; ++i) {
    const size_t p = f.find_last_of('/');
    f = (p == std::string::npos) ? std::string(".") : f.substr(0, p);
  }
  return f;
}
void set_yaml_dir(const std::string& dir) { g_yaml_dir = dir; }
void set_robot_data_dir(const std::string& dir) { g_robot_dir = dir; }
static std::string yaml_dir() {
  if (!g_yaml_dir.empty()) return g_yaml_dir;
  if (const char* e = std::getenv("EAGLE_MPC_YAML_DIR")) return e;
  return repo_root() + "/yaml";
}
static std::string robot_dir() {
  if (!g_robot_dir.empty()) return g_robot_dir;
  if (const char* e = std::getenv("EAGLE_MPC_ROBOT_DATA_DIR")) return e;
  return repo_root() + "/fixtures/urdf";
}
std::string getYamlPath(const std::string& p) { return p.find("/", 0) == 0 ? p : yaml_dir() + "/" + p; }
std::string getUrdfPath(const std::string& p) { return p.find("/", 0) == 0 ? p : robot_dir() + "/" + p; }

// ---- converters -----------------------------------------------------------------------------------------------------
static std::string strip_spaces(const std::string& s) {
  std::string o;
  for (char c : s) if (!std::isspace((unsigned char)c)) o.push_back(c);
  return o;
}
static std::vector<std::string> split_top_level(const std::string& inner) {
  std::vector<std::string> out;
  std::string cur;
  int depth = 0;
  for (char c : inner) {
    if (c == '[' || c == '{') depth++;
    if (c == ']' || c == '}') depth--;
    if (c == ',' && depth == 0) { if (!cur.empty()) out.push_back(cur); cur.clear(); }
    else cur.push_back(c);
  }
  if (!cur.empty()) out.push_back(cur);
  return out;
}
template <> int converter<int>::convert(const std::string& v) { return std::stoi(v); }
template <> unsigned int converter<unsigned int>::convert(const std::string& v) { return (unsigned int)std::stoul(v); }
template <> double converter<double>::convert(const std::string& v) { return std::stod(v); }
template <> std::string converter<std::string>::convert(const std::string& v) { return v; }
template <> bool converter<bool>::convert(const std::string& v) {
  if (v == "true") return true;
  if (v == "false") return false;
  throw std::runtime_error("Invalid conversion to bool (Must be either \"true\" or \"false\"). String provided: " + v);
}
template <> VectorXd converter<VectorXd>::convert(const std::string& val) {
  const std::string s = strip_spaces(val);
  // splitMatrixStringRepresentation: "[num(,num)*]" with num = -?[0-9]*(\.[0-9]+)? (no exponent notation)
  static const std::regex rgx("\\[((?:(?:-?[0-9]*)(?:\\.[0-9]+)?,?)+)\\]");
  std::smatch m;
  if (!std::regex_match(s, m, rgx))
    throw std::runtime_error("Invalid string representation of a Matrix. Correct format is [([num,num],)?(num(,num)*)?]. String provided: " + val);
  VectorXd out;
  for (const std::string& tok : split_top_level(m[1].str())) out.push_back(std::stod(tok));
  return out;
}
template <> std::vector<std::string> converter<std::vector<std::string>>::convert(const std::string& val) {
  const std::string s = strip_spaces(val);
  if (s.size() < 2 || s.front() != '[' || s.back() != ']')
    throw std::runtime_error("Invalid string format representing a list-like structure. Correct format is [(value)?(,value)*]. String provided: " + val);
  return split_top_level(s.substr(1, s.size() - 2));
}
// "{k:v,k:[..]}" -> map
template <> std::map<std::string, std::string> converter<std::map<std::string, std::string>>::convert(const std::string& val) {
  const std::string s = strip_spaces(val);
  if (s.size() < 2 || s.front() != '{' || s.back() != '}')
    throw std::runtime_error("Invalid string representation of a Map. String provided: " + val);
  std::map<std::string, std::string> out;
  for (const std::string& kv : split_top_level(s.substr(1, s.size() - 2))) {
    const size_t c = kv.find(':');
    if (c == std::string::npos) throw std::runtime_error("Invalid map entry: " + kv);
    out[kv.substr(0, c)] = kv.substr(c + 1);
  }
  return out;
}
template <> std::vector<std::map<std::string, std::string>> converter<std::vector<std::map<std::string, std::string>>>::convert(const std::string& val) {
  std::vector<std::map<std::string, std::string>> out;
  for (const std::string& item : converter<std::vector<std::string>>::convert(val))
    out.push_back(converter<std::map<std::string, std::string>>::convert(item));
  return out;
}

// ---- YAML subset ------------------------------------------------------------------------------------------------------
namespace {
struct Node {
  enum Type { Undefined, Scalar, Sequence, Map } type = Undefined;
  std::string scalar;
  std::vector<Node> seq;
  std::vector<std::pair<std::string, Node>> map;
  const Node* get(const std::string& k) const {
    for (auto& kv : map) if (kv.first == k) return &kv.second;
    return nullptr;
  }
};
struct Line { int indent; std::string text; };

std::string rtrim(const std::string& s) {
  size_t e = s.size();
  while (e > 0 && std::isspace((unsigned char)s[e - 1])) --e;
  return s.substr(0, e);
}
std::string ltrim(const std::string& s) {
  size_t b = 0;
  while (b < s.size() && std::isspace((unsigned char)s[b])) ++b;
  return s.substr(b);
}
std::string unquote(const std::string& s) {
  std::string t = rtrim(ltrim(s));
  if (t.size() >= 2 && ((t.front() == '"' && t.back() == '"') || (t.front() == '\'' && t.back() == '\''))) return t.substr(1, t.size() - 2);
  return t;
}
std::string strip_comment(const std::string& s) {
  bool inq = false; char q = 0;
  for (size_t i = 0; i < s.size(); ++i) {
    const char c = s[i];
    if (inq) { if (c == q) inq = false; }
    else if (c == '"' || c == '\'') { inq = true; q = c; }
    else if (c == '#' && (i == 0 || std::isspace((unsigned char)s[i - 1]))) return s.substr(0, i);
  }
  return s;
}
int bracket_balance(const std::string& s) {
  int d = 0;
  for (char c : s) { if (c == '[') d++; if (c == ']') d--; }
  return d;
}
Node parse_flow_seq(const std::string& text) {
  Node n; n.type = Node::Sequence;
  std::string s = rtrim(ltrim(text));
  s = s.substr(1, s.size() - 2);
  for (const std::string& tok : split_top_level(s)) {
    const std::string t = unquote(tok);
    if (t.empty()) continue;  // trailing comma
    Node e;
    if (t.front() == '[') e = parse_flow_seq(t);
    else { e.type = Node::Scalar; e.scalar = t; }
    n.seq.push_back(e);
  }
  return n;
}
Node parse_value_text(const std::string& v) {
  const std::string t = rtrim(ltrim(v));
  if (!t.empty() && t.front() == '[') return parse_flow_seq(t);
  Node n; n.type = Node::Scalar; n.scalar = unquote(t);
  return n;
}

struct YamlParser {
  std::vector<Line> lines;
  size_t pos = 0;
  explicit YamlParser(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("Couldn't load file: " + path);
    std::string raw;
    while (std::getline(f, raw)) {
      std::string s = rtrim(strip_comment(raw));
      if (ltrim(s).empty()) continue;
      int ind = 0;
      while (ind < (int)s.size() && s[ind] == ' ') ++ind;
      lines.push_back({ind, s.substr(ind)});
    }
    // join flow sequences continued over several lines
    std::vector<Line> joined;
    for (size_t i = 0; i < lines.size(); ++i) {
      Line l = lines[i];
      int bal = bracket_balance(l.text);
      while (bal > 0 && i + 1 < lines.size()) { ++i; l.text += " " + lines[i].text; bal = bracket_balance(l.text); }
      joined.push_back(l);
    }
    lines.swap(joined);
  }
  // split "key: value" at the first ':' followed by space/end (outside quotes/brackets)
  static bool split_key(const std::string& t, std::string& key, std::string& val) {
    bool inq = false; char q = 0; int depth = 0;
    for (size_t i = 0; i < t.size(); ++i) {
      const char c = t[i];
      if (inq) { if (c == q) inq = false; continue; }
      if (c == '"' || c == '\'') { inq = true; q = c; continue; }
      if (c == '[') depth++;
      if (c == ']') depth--;
      if (c == ':' && depth == 0 && (i + 1 == t.size() || t[i + 1] == ' ')) {
        key = unquote(t.substr(0, i));
        val = (i + 1 < t.size()) ? ltrim(t.substr(i + 1)) : "";
        return true;
      }
    }
    return false;
  }
  Node parse_block(int indent) {
    Node n;
    if (pos >= lines.size()) return n;
    if (lines[pos].text.rfind("- ", 0) == 0 || lines[pos].text == "-") {
      n.type = Node::Sequence;
      while (pos < lines.size() && lines[pos].indent == indent && (lines[pos].text.rfind("- ", 0) == 0 || lines[pos].text == "-")) {
        const std::string rest = lines[pos].text.size() > 2 ? ltrim(lines[pos].text.substr(2)) : "";
        const int item_indent = indent + (int)(lines[pos].text.size() - ltrim(lines[pos].text.substr(1)).size());
        std::string k, v;
        if (rest.empty()) { ++pos; n.seq.push_back(parse_block(pos < lines.size() ? lines[pos].indent : indent + 2)); }
        else if (rest.front() != '[' && split_key(rest, k, v)) {
          // a map starting on the dash line: rewrite the line as a map entry at item_indent
          lines[pos].indent = item_indent; lines[pos].text = rest;
          n.seq.push_back(parse_block(item_indent));
        } else { n.seq.push_back(parse_value_text(rest)); ++pos; }
      }
      return n;
    }
    n.type = Node::Map;
    while (pos < lines.size() && lines[pos].indent == indent && lines[pos].text.rfind("- ", 0) != 0) {
      std::string k, v;
      if (!split_key(lines[pos].text, k, v)) throw std::runtime_error("YAML: cannot parse line '" + lines[pos].text + "'");
      ++pos;
      Node child;
      if (!v.empty()) child = parse_value_text(v);
      else if (pos < lines.size() && lines[pos].indent > indent) {
        if (lines[pos].text.front() == '[') { child = parse_flow_seq(lines[pos].text); ++pos; }  // value on its own line
        else child = parse_block(lines[pos].indent);
      } else if (pos < lines.size() && lines[pos].indent == indent && lines[pos].text.rfind("- ", 0) == 0) {
        child = parse_block(indent);  // sequence at the same indentation as its key
      }
      n.map.push_back({k, child});
    }
    return n;
  }
  Node parse() { return lines.empty() ? Node() : parse_block(lines[0].indent); }
};

Node load_yaml(const std::string& path) { YamlParser p(path); return p.parse(); }

std::string atomic_string(const Node& n) {  // parseAtomicNode
  switch (n.type) {
    case Node::Scalar: return n.scalar;
    case Node::Sequence: {
      std::string s = "[";
      for (size_t i = 0; i < n.seq.size(); ++i) s += (i ? "," : "") + atomic_string(n.seq[i]);
      return s + "]";
    }
    case Node::Map: {
      std::string s = "{";
      for (size_t i = 0; i < n.map.size(); ++i) s += (i ? "," : "") + n.map[i].first + ":" + atomic_string(n.map[i].second);
      return s + "}";
    }
    default: return "";
  }
}
bool seq_is_atomic(const Node& n) {  // isAtomic("", sequence): atomic unless it contains maps with non-$ keys
  for (const Node& it : n.seq) {
    if (it.type == Node::Map) {
      for (auto& kv : it.map) if (kv.first.empty() || kv.first[0] != '$') return false;
    } else if (it.type == Node::Sequence && !seq_is_atomic(it)) return false;
  }
  return true;
}

struct Flattener {
  std::map<std::string, std::string>& params;
  void insert(std::string key, const std::string& value) {
    if (!key.empty() && key[0] == '/') key = key.substr(1);
    params.insert({key, value});  // duplicates are skipped, as insertRegister does
  }
  void walk(const Node& node, const std::string& name) {
    switch (node.type) {
      case Node::Scalar: insert(name, node.scalar); break;
      case Node::Sequence:
        if (seq_is_atomic(node)) insert(name, atomic_string(node));
        else for (const Node& it : node.seq) walk(it, name);
        break;
      case Node::Map:
        for (auto& kv : node.map) {
          if (!kv.first.empty() && kv.first[0] == '$') insert(name + "/" + kv.first.substr(1), atomic_string(kv.second));
          else if (kv.first == "follow") walk(load_yaml(getYamlPath(kv.second.scalar)), name);
          else walk(kv.second, name + "/" + kv.first);
        }
        break;
      default: break;
    }
  }
};
}  // namespace

ParserYaml::ParserYaml(const std::string& file, const std::string& path_root, bool freely_parse) {
  std::string path = file;
  if (!(file.size() && file[0] == '/') && !path_root.empty()) path = path_root + (path_root.back() == '/' ? "" : "/") + file;
  const Node root = load_yaml(path);
  Flattener fl{params_};
  if (freely_parse) { fl.walk(root, ""); return; }
  const Node* traj = root.get("trajectory");
  if (!traj || traj->type != Node::Map) {
    const Node* mpc = root.get("mpc_controller");
    if (!mpc || mpc->type != Node::Map)
      throw std::runtime_error("Could not find neither a trajectory or an mpc_controller node. Please make sure that your YAML file " + path + " starts with 'trajectory:' or 'mpc_controller:'");
    const Node* robot = mpc->get("robot");
    if (!robot || robot->type != Node::Map)
      throw std::runtime_error("Could not find robot node. Please make sure that the 'trajectory' node in YAML file " + path + " has a 'robot' entry");
    for (auto& kv : mpc->map)
      if (kv.first != "robot") fl.insert("mpc_controller/" + kv.first, atomic_string(kv.second));
    fl.walk(*robot, "robot");
    return;
  }
  const Node* robot = traj->get("robot");
  if (!robot || robot->type != Node::Map)
    throw std::runtime_error("Could not find robot node. Please make sure that the 'trajectory' node in YAML file " + path + " has a 'robot' entry");
  if (const Node* init = traj->get("initial_state"))
    if (init->type == Node::Sequence) fl.insert("initial_state", atomic_string(*init));
  std::string stages_str = "[";
  const Node* stages = traj->get("stages");
  if (!stages || stages->type != Node::Sequence)
    throw std::runtime_error("Error parsing stages @" + path + ". Make sure every stage has a name, duration and, at least, one cost.");
  bool first = true;
  for (const Node& st : stages->seq) {
    const Node *nm = st.get("name"), *du = st.get("duration"), *costs = st.get("costs");
    if (!nm || !du || !costs || costs->type != Node::Sequence)
      throw std::runtime_error("Error parsing stages @" + path + ". Make sure every stage has a name, duration and, at least, one cost.");
    const std::string transition = st.get("transition") ? "true" : "false";  // presence, not value (:274-278)
    // every cost / contact entry needs a name: the reference's try/catch turns the yaml-cpp exception into this message
    auto name_of = [&](const Node& entry) -> const std::string& {
      const Node* n = entry.get("name");
      if (!n) throw std::runtime_error("Error parsing stages @" + path + ". Make sure every stage has a name, duration and, at least, one cost.");
      return n->scalar;
    };
    std::string clist = "[";
    for (size_t i = 0; i < costs->seq.size(); ++i) clist += (i ? "," : "") + name_of(costs->seq[i]);
    clist += "]";
    std::string entry = "{name:" + nm->scalar + ",duration:" + du->scalar + ",transition:" + transition + ",costs:" + clist;
    if (const Node* contacts = st.get("contacts")) {
      std::string l = "[";
      for (size_t i = 0; i < contacts->seq.size(); ++i) l += (i ? "," : "") + name_of(contacts->seq[i]);
      entry += ",contacts:" + l + "]";
    }
    entry += "}";
    stages_str += (first ? "" : ",") + entry;
    first = false;
  }
  fl.insert("stages", stages_str + "]");
  fl.walk(*robot, "robot");
  if (const Node* pp = traj->get("problem_params"))
    if (pp->type == Node::Map) fl.walk(*pp, "problem_params");
  for (const Node& st : stages->seq) {
    const std::string base = "stages/" + st.get("name")->scalar;
    fl.insert(base + "/name", st.get("name")->scalar);
    fl.insert(base + "/duration", st.get("duration")->scalar);
    fl.insert(base + "/transition", st.get("transition") ? "true" : "false");
    for (const Node& c : st.get("costs")->seq) fl.walk(c, base + "/costs/" + c.get("name")->scalar);
    if (const Node* contacts = st.get("contacts"))
      for (const Node& c : contacts->seq) fl.walk(c, base + "/contacts/" + c.get("name")->scalar);
  }
}

}  // namespace eagle_mpc

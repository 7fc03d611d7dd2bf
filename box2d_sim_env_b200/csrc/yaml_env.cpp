// YAML environment loader of the batched backend (host only, setup time).
// Replaces the yaml-cpp decoders of the reference (include/sim_env/Box2DIOUtils.h:83-305,
// src/sim_env/Box2DIOUtils.cpp:16-92).  yaml-cpp is not available, so this is a small reader for the
// YAML subset the environment files use: block mappings and sequences, flow sequences of scalars,
// plain scalars, comments.
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>

#include "desc.h"

namespace b2g {
namespace {

struct Node {
    enum Kind { Null, Scalar, Seq, Map } kind = Null;
    std::string scalar;
    std::vector<Node> seq;
    std::vector<std::pair<std::string, Node>> map;

    const Node* find(const std::string& key) const {
        for (auto& kv : map)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    const Node& at(const std::string& key, const std::string& ctx) const {
        const Node* n = find(key);
        if (!n) throw std::runtime_error("missing key '" + key + "' in " + ctx);
        return *n;
    }
};

struct Line {
    int indent;
    std::string text;
    int number;
};

std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}

std::string stripComment(const std::string& s) {
    bool inS = false, inD = false;
    for (size_t i = 0; i < s.size(); ++i) {
        char c = s[i];
        if (c == '\'' && !inD) inS = !inS;
        else if (c == '"' && !inS) inD = !inD;
        else if (c == '#' && !inS && !inD && (i == 0 || std::isspace((unsigned char)s[i - 1]))) return s.substr(0, i);
    }
    return s;
}

std::string unquote(const std::string& s) {
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\'')))
        return s.substr(1, s.size() - 2);
    return s;
}

Node parseInline(const std::string& textIn, int lineNo) {
    std::string text = trim(textIn);
    Node n;
    if (text.empty() || text == "~" || text == "null") return n;
    if (text.front() == '[') {
        if (text.back() != ']') throw std::runtime_error("line " + std::to_string(lineNo) + ": unterminated flow sequence");
        n.kind = Node::Seq;
        std::string inner = text.substr(1, text.size() - 2);
        std::string item;
        int depth = 0;
        for (char c : inner) {
            if (c == '[') ++depth;
            if (c == ']') --depth;
            if (c == ',' && depth == 0) {
                if (!trim(item).empty()) n.seq.push_back(parseInline(item, lineNo));
                item.clear();
            } else {
                item.push_back(c);
            }
        }
        if (!trim(item).empty()) n.seq.push_back(parseInline(item, lineNo));
        return n;
    }
    n.kind = Node::Scalar;
    n.scalar = unquote(text);
    return n;
}

// position of the ':' that separates a mapping key from its value, or npos
size_t keyColon(const std::string& t) {
    if (t.empty() || t[0] == '[' || t[0] == '"' || t[0] == '\'') return std::string::npos;
    for (size_t i = 0; i < t.size(); ++i) {
        if (t[i] == ':' && (i + 1 == t.size() || t[i + 1] == ' ')) return i;
        if (t[i] == '[' || t[i] == ',') return std::string::npos;
    }
    return std::string::npos;
}

struct Parser {
    std::vector<Line> lines;
    size_t pos = 0;

    Node parseBlock(int indent) {
        Node n;
        if (pos >= lines.size() || lines[pos].indent < indent) return n;
        int blockIndent = lines[pos].indent;
        if (lines[pos].text[0] == '-' && (lines[pos].text.size() == 1 || lines[pos].text[1] == ' ')) {
            n.kind = Node::Seq;
            while (pos < lines.size() && lines[pos].indent == blockIndent && lines[pos].text[0] == '-' &&
                   (lines[pos].text.size() == 1 || lines[pos].text[1] == ' ')) {
                std::string rest = lines[pos].text.substr(1);
                size_t lead = 0;
                while (lead < rest.size() && rest[lead] == ' ') ++lead;
                rest = rest.substr(lead);
                if (rest.empty()) {
                    ++pos;
                    n.seq.push_back(parseBlock(blockIndent + 1));
                } else if (keyColon(rest) != std::string::npos) {
                    // "- key: value": re-interpret the line as the first key of a mapping indented past the dash
                    lines[pos].indent = blockIndent + 1 + (int)lead;
                    lines[pos].text = rest;
                    n.seq.push_back(parseBlock(lines[pos].indent));
                } else {
                    n.seq.push_back(parseInline(rest, lines[pos].number));
                    ++pos;
                }
            }
            return n;
        }
        n.kind = Node::Map;
        while (pos < lines.size() && lines[pos].indent == blockIndent) {
            const std::string& t = lines[pos].text;
            if (t[0] == '-' && (t.size() == 1 || t[1] == ' ')) break;
            size_t c = keyColon(t);
            if (c == std::string::npos)
                throw std::runtime_error("line " + std::to_string(lines[pos].number) + ": expected 'key: value'");
            std::string key = unquote(trim(t.substr(0, c)));
            std::string val = trim(t.substr(c + 1));
            int lineNo = lines[pos].number;
            ++pos;
            if (!val.empty()) {
                n.map.emplace_back(key, parseInline(val, lineNo));
            } else if (pos < lines.size() && (lines[pos].indent > blockIndent ||
                                              (lines[pos].indent == blockIndent && lines[pos].text[0] == '-' &&
                                               (lines[pos].text.size() == 1 || lines[pos].text[1] == ' ')))) {
                n.map.emplace_back(key, parseBlock(lines[pos].indent));
            } else {
                n.map.emplace_back(key, Node());
            }
        }
        return n;
    }
};

Node loadFile(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Could not open the file " + path + " as it does not exist.");
    Parser p;
    std::string raw;
    int number = 0;
    auto flowDepth = [](const std::string& t) {  // unbalanced '[' of a line (quotes respected)
        int depth = 0;
        bool inS = false, inD = false;
        for (char ch : t) {
            if (ch == '\'' && !inD) inS = !inS;
            else if (ch == '"' && !inS) inD = !inD;
            else if (!inS && !inD && ch == '[') ++depth;
            else if (!inS && !inD && ch == ']') --depth;
        }
        return depth;
    };
    while (std::getline(in, raw)) {
        ++number;
        if (!raw.empty() && raw.back() == '\r') raw.pop_back();
        std::string s = stripComment(raw);
        // a flow sequence may run over several lines ("- [0.0, 0.0,\n   0.8, 0.0]"): join them
        for (int depth = flowDepth(s); depth > 0;) {
            std::string more;
            if (!std::getline(in, more)) throw std::runtime_error(path + ":" + std::to_string(number) + ": unterminated flow sequence");
            ++number;
            if (!more.empty() && more.back() == '\r') more.pop_back();
            more = stripComment(more);
            depth += flowDepth(more);
            s += " " + trim(more);
        }
        size_t ind = 0;
        while (ind < s.size() && s[ind] == ' ') ++ind;
        std::string t = trim(s);
        if (t.empty() || t == "---") continue;
        if (s.find('\t') != std::string::npos && s.find('\t') < ind + 1)
            throw std::runtime_error(path + ":" + std::to_string(number) + ": tab indentation is not supported");
        p.lines.push_back(Line{(int)ind, t, number});
    }
    try {
        return p.parseBlock(0);
    } catch (const std::exception& e) {
        throw std::runtime_error(path + ": " + e.what());
    }
}

float asFloat(const Node& n, const std::string& ctx) {
    if (n.kind != Node::Scalar) throw std::runtime_error(ctx + ": expected a number");
    char* end = nullptr;
    double v = std::strtod(n.scalar.c_str(), &end);
    if (end == n.scalar.c_str() || *end != 0) throw std::runtime_error(ctx + ": '" + n.scalar + "' is not a number");
    return (float)v;
}
bool asBool(const Node& n, const std::string& ctx) {
    if (n.kind != Node::Scalar) throw std::runtime_error(ctx + ": expected a boolean");
    std::string s;
    for (char c : n.scalar) s.push_back((char)std::tolower((unsigned char)c));
    if (s == "true" || s == "yes" || s == "on" || s == "y") return true;
    if (s == "false" || s == "no" || s == "off" || s == "n") return false;
    throw std::runtime_error(ctx + ": '" + n.scalar + "' is not a boolean");
}
std::string asString(const Node& n, const std::string& ctx) {
    if (n.kind != Node::Scalar) throw std::runtime_error(ctx + ": expected a string");
    return n.scalar;
}
std::vector<float> asFloats(const Node& n, const std::string& ctx, int expect = -1) {
    if (n.kind != Node::Seq) throw std::runtime_error(ctx + ": expected a sequence");
    std::vector<float> out;
    for (auto& e : n.seq) out.push_back(asFloat(e, ctx));
    if (expect >= 0 && (int)out.size() != expect)
        throw std::runtime_error(ctx + ": expected " + std::to_string(expect) + " values");
    return out;
}
void copy2(float dst[2], const std::vector<float>& v) { dst[0] = v[0]; dst[1] = v[1]; }

LinkDescription decodeLink(const Node& node, EnvironmentDescription& ed, const std::string& file) {
    // Box2DIOUtils.h:127-145
    LinkDescription ld;
    std::string ctx = file + " link";
    ld.name = asString(node.at("name", ctx), ctx);
    ctx += " " + ld.name;
    ld.mass = asFloat(node.at("mass", ctx), ctx + " mass");
    ld.contact_friction = asFloat(node.at("contact_friction", ctx), ctx + " contact_friction");
    ld.restitution = asFloat(node.at("restitution", ctx), ctx + " restitution");
    // The decoder reads ground_friction / ground_torque_integral; the shipped data still says
    // trans_friction / rot_friction.  Accept both, prefer the decoder's names.
    if (const Node* n = node.find("ground_torque_integral")) {
        ld.ground_torque_integral = asFloat(*n, ctx + " ground_torque_integral");
    } else if (const Node* n2 = node.find("rot_friction")) {
        ld.ground_torque_integral = asFloat(*n2, ctx + " rot_friction");
        ed.warnings.push_back(ctx + ": stale key rot_friction used as ground_torque_integral");
    } else {
        throw std::runtime_error("missing key 'ground_torque_integral' in " + ctx);
    }
    if (const Node* n = node.find("ground_friction")) {
        ld.ground_friction = asFloat(*n, ctx + " ground_friction");
    } else if (const Node* n2 = node.find("trans_friction")) {
        ld.ground_friction = asFloat(*n2, ctx + " trans_friction");
        ed.warnings.push_back(ctx + ": stale key trans_friction used as ground_friction");
    } else {
        throw std::runtime_error("missing key 'ground_friction' in " + ctx);
    }
    const Node& geom = node.at("geometry", ctx);
    if (geom.kind != Node::Seq) throw std::runtime_error(ctx + ": geometry must be a sequence of polygons");
    for (auto& poly : geom.seq) {
        std::vector<float> p = asFloats(poly, ctx + " geometry");
        if (p.size() % 2 != 0 || p.size() < 6)
            throw std::runtime_error(ctx + ": a polygon needs an even number (>= 6) of floats");
        ld.polygons.push_back(p);
    }
    // optional ball approximation (Box2DIOUtils.h:139-142): a sequence of yaml-cpp pairs [[x, y, z], radius]
    if (const Node* balls = node.find("balls")) {
        if (balls->kind != Node::Seq && balls->kind != Node::Null)
            throw std::runtime_error(ctx + ": balls must be a sequence of [[x, y, z], radius] pairs");
        for (auto& b : balls->seq) {
            if (b.kind != Node::Seq || b.seq.size() != 2)
                throw std::runtime_error(ctx + ": a ball is a pair [[x, y, z], radius]");
            std::vector<float> c = asFloats(b.seq[0], ctx + " ball centre", 3);
            ld.balls.push_back({c[0], c[1], c[2], asFloat(b.seq[1], ctx + " ball radius")});
        }
    }
    return ld;
}

JointDescription decodeJoint(const Node& node, const std::string& file) {
    // Box2DIOUtils.h:167-182
    JointDescription jd;
    std::string ctx = file + " joint";
    jd.name = asString(node.at("name", ctx), ctx);
    ctx += " " + jd.name;
    jd.joint_type = asString(node.at("joint_type", ctx), ctx);
    jd.axis_orientation = asFloat(node.at("axis_orientation", ctx), ctx);
    jd.link_a = asString(node.at("link_a", ctx), ctx);
    jd.link_b = asString(node.at("link_b", ctx), ctx);
    copy2(jd.axis, asFloats(node.at("axis", ctx), ctx + " axis", 2));
    copy2(jd.position_limits, asFloats(node.at("position_limits", ctx), ctx + " position_limits", 2));
    copy2(jd.velocity_limits, asFloats(node.at("velocity_limits", ctx), ctx + " velocity_limits", 2));
    copy2(jd.acceleration_limits, asFloats(node.at("acceleration_limits", ctx), ctx + " acceleration_limits", 2));
    jd.actuated = asBool(node.at("actuated", ctx), ctx + " actuated");
    return jd;
}

ObjectDescription decodeObject(const Node& node, EnvironmentDescription& ed, const std::string& file) {
    // Box2DIOUtils.h:203-257
    ObjectDescription od;
    if (node.kind != Node::Map) throw std::runtime_error(file + ": object file must be a mapping");
    od.name = asString(node.at("name", file), file);
    const Node* links = node.find("links");
    if (links && links->kind == Node::Seq)
        for (auto& l : links->seq) od.links.push_back(decodeLink(l, ed, file));
    const Node* joints = node.find("joints");
    if (joints && joints->kind == Node::Seq)
        for (auto& j : joints->seq) od.joints.push_back(decodeJoint(j, file));
    std::string msg;
    if (!verifyKinematicStructure(od, msg))
        throw std::runtime_error("The Kinematic structure of object " + od.name + " is invalid. " + msg);
    return od;
}

std::string resolveFileName(const std::string& file, const std::string& root) {
    // Box2DIOUtils.cpp:6-14
    if (!file.empty() && file[0] == '/') return file;
    if (root.empty()) return file;
    return root + "/" + file;
}

}  // namespace

bool verifyKinematicStructure(ObjectDescription& od, std::string& message) {
    std::map<std::string, int> numParents;
    for (auto& l : od.links) numParents[l.name] = 0;
    for (auto& j : od.joints) {
        if (!numParents.count(j.link_a)) {
            message = "Parent link \"" + j.link_a + "\" of joint \"" + j.name + "\" is unknown.";
            return false;
        }
        auto it = numParents.find(j.link_b);
        if (it == numParents.end()) {
            message = "Child link \"" + j.link_b + "\" of joint \"" + j.name + "\" is unknown.";
            return false;
        }
        it->second += 1;
    }
    bool foundRoot = false;
    for (auto& l : od.links) {
        if (numParents[l.name] == 0) {
            if (foundRoot) {
                message = "Found an additional link with no parent: \"" + l.name + "\"; previous root is \"" + od.base_link + "\"";
                return false;
            }
            foundRoot = true;
            od.base_link = l.name;
        } else if (numParents[l.name] > 1) {
            message = "Link \"" + l.name + "\" has more than one parent joint.";
            return false;
        }
    }
    if (!foundRoot && !od.links.empty()) {
        message = "No root link (kinematic loop).";
        return false;
    }
    return true;
}

void parseYAML(const std::string& filename, EnvironmentDescription& ed) {
    // Box2DIOUtils.cpp:63-92
    std::string root;
    size_t slash = filename.find_last_of('/');
    if (slash != std::string::npos) root = filename.substr(0, slash);
    Node node = loadFile(filename);
    if (node.kind != Node::Map) throw std::runtime_error(filename + ": environment file must be a mapping");
    if (const Node* s = node.find("scale")) ed.scale = asFloat(*s, filename + " scale");
    else { ed.scale = 1.0f; ed.warnings.push_back("No scale specified. Using default scale (1.0)."); }
    if (const Node* wb = node.find("world_bounds")) {
        std::vector<float> v = asFloats(*wb, filename + " world_bounds", 4);
        for (int i = 0; i < 4; ++i) ed.world_bounds[i] = v[i];
    } else {
        for (int i = 0; i < 4; ++i) ed.world_bounds[i] = 0.0f;
        ed.warnings.push_back("No world bounds specified. Using default world bounds.");
    }
    if (const Node* robots = node.find("robots")) {  // Box2DIOUtils.cpp:40-61
        for (auto& entry : robots->seq) {
            std::string name;
            if (const Node* n = entry.find("name")) name = asString(*n, filename + " robot name");
            std::string file = resolveFileName(asString(entry.at("filename", filename + " robots"), filename), root);
            Node rnode = loadFile(file);
            ObjectDescription od = decodeObject(rnode, ed, file);
            od.is_robot = true;
            if (const Node* ba = rnode.find("base_actuation")) {  // Box2DIOUtils.h:285-303
                od.is_static = false;
                std::string ctx = file + " base_actuation";
                copy2(od.rotational_acceleration_limits, asFloats(ba->at("rotational_acceleration_limits", ctx), ctx, 2));
                copy2(od.rotational_velocity_limits, asFloats(ba->at("rotational_velocity_limits", ctx), ctx, 2));
                copy2(od.translational_acceleration_limits, asFloats(ba->at("translational_acceleration_limits", ctx), ctx, 2));
                copy2(od.translational_velocity_limits, asFloats(ba->at("translational_velocity_limits", ctx), ctx, 2));
                od.use_center_of_mass = asBool(ba->at("use_center_of_mass", ctx), ctx);
            } else {
                od.is_static = true;
            }
            if (!name.empty()) od.name = name;
            ed.robots.push_back(od);
        }
    }
    if (const Node* objects = node.find("objects")) {  // Box2DIOUtils.cpp:16-38
        for (auto& entry : objects->seq) {
            std::string name;
            if (const Node* n = entry.find("name")) name = asString(*n, filename + " object name");
            std::string file = resolveFileName(asString(entry.at("filename", filename + " objects"), filename), root);
            Node onode = loadFile(file);
            ObjectDescription od = decodeObject(onode, ed, file);
            if (!name.empty()) od.name = name;
            if (const Node* st = entry.find("static")) od.is_static = asBool(*st, filename + " static");
            else od.is_static = false;
            ed.objects.push_back(od);
        }
    }
    if (const Node* states = node.find("states")) {
        for (auto& s : states->seq) {
            std::string name = asString(s.at("name", filename + " states"), filename);
            const Node& st = s.at("state", filename + " states " + name);
            StateDescription sd;
            sd.configuration = asFloats(st.at("configuration", name), name + " configuration");
            sd.velocity = asFloats(st.at("velocity", name), name + " velocity");
            ed.states[name] = sd;
        }
    }
}

}  // namespace b2g

// Host model builder (setup time).  Re-derives what Box2DWorld::createWorld and the Box2DLink / Box2DJoint /
// Box2DObject / Box2DRobot constructors leave behind in Box2D (src/sim_env/Box2DWorld.cpp:44-148, 872-948,
// 1292-1363, 2046-2067, 3644-3721), using Box2D 2.3.1's published shape arithmetic (b2PolygonShape::Set,
// ComputeMass; b2Body::ResetMassData), and flattens it into the word blob described in model.h.
#include "model_build.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <stdexcept>

namespace b2g {
namespace {

struct P2 {
    float x, y;
};
inline P2 sub(P2 a, P2 b) { return P2{a.x - b.x, a.y - b.y}; }
inline float cross2(P2 a, P2 b) { return a.x * b.y - a.y * b.x; }
inline float dot2(P2 a, P2 b) { return a.x * b.x + a.y * b.y; }

constexpr float kLinearSlop = 0.005f;
constexpr float kPolygonRadius = 2.0f * kLinearSlop;
constexpr float kAabbExtension = 0.1f;
constexpr int kMaxPolyVerts = 8;

struct Hull {
    int n = 0;
    P2 v[kMaxPolyVerts];
    P2 nrm[kMaxPolyVerts];
    P2 centroid{0.0f, 0.0f};
};

void boxHull(Hull& h, float hx, float hy) {
    h.n = 4;
    h.v[0] = P2{-hx, -hy}; h.v[1] = P2{hx, -hy}; h.v[2] = P2{hx, hy}; h.v[3] = P2{-hx, hy};
    h.nrm[0] = P2{0.0f, -1.0f}; h.nrm[1] = P2{1.0f, 0.0f}; h.nrm[2] = P2{0.0f, 1.0f}; h.nrm[3] = P2{-1.0f, 0.0f};
    h.centroid = P2{0.0f, 0.0f};
}

// b2PolygonShape::Set (2.3.1): weld, gift-wrap from the right-most point, normals, centroid
void convexHull(Hull& h, const std::vector<P2>& in) {
    int n = std::min((int)in.size(), kMaxPolyVerts);
    if (n < 3) { boxHull(h, 1.0f, 1.0f); return; }
    P2 pts[kMaxPolyVerts];
    int np = 0;
    for (int i = 0; i < n; ++i) {
        bool keep = true;
        for (int j = 0; j < np && keep; ++j) {
            P2 d = sub(in[i], pts[j]);
            if (dot2(d, d) < 0.5f * kLinearSlop) keep = false;
        }
        if (keep) pts[np++] = in[i];
    }
    if (np < 3) { boxHull(h, 1.0f, 1.0f); return; }
    int start = 0;
    float bestX = pts[0].x;
    for (int i = 1; i < np; ++i) {
        float x = pts[i].x;
        if (x > bestX || (x == bestX && pts[i].y < pts[start].y)) { start = i; bestX = x; }
    }
    int order[kMaxPolyVerts];
    int m = 0;
    int cur = start;
    while (true) {
        order[m] = cur;
        int next = 0;
        for (int j = 1; j < np; ++j) {
            if (next == cur) { next = j; continue; }
            P2 r = sub(pts[next], pts[order[m]]);
            P2 v = sub(pts[j], pts[order[m]]);
            float c = cross2(r, v);
            if (c < 0.0f) next = j;
            if (c == 0.0f && dot2(v, v) > dot2(r, r)) next = j;
        }
        ++m;
        cur = next;
        if (next == start) break;
    }
    if (m < 3) { boxHull(h, 1.0f, 1.0f); return; }
    h.n = m;
    for (int i = 0; i < m; ++i) h.v[i] = pts[order[i]];
    for (int i = 0; i < m; ++i) {
        P2 e = sub(h.v[(i + 1 < m) ? i + 1 : 0], h.v[i]);
        P2 nr{1.0f * e.y, -1.0f * e.x};  // b2Cross(edge, 1.0f)
        float len = sqrtf(nr.x * nr.x + nr.y * nr.y);
        if (!(len < FLT_EPSILON)) {
            float inv = 1.0f / len;
            nr.x *= inv;
            nr.y *= inv;
        }
        h.nrm[i] = nr;
    }
    // centroid by triangle fan about the origin
    P2 c{0.0f, 0.0f};
    float area = 0.0f;
    const float third = 1.0f / 3.0f;
    for (int i = 0; i < m; ++i) {
        P2 p1{0.0f, 0.0f};
        P2 p2 = h.v[i];
        P2 p3 = (i + 1 < m) ? h.v[i + 1] : h.v[0];
        float D = cross2(sub(p2, p1), sub(p3, p1));
        float tri = 0.5f * D;
        area += tri;
        float wgt = tri * third;
        c.x += wgt * (p1.x + p2.x + p3.x);
        c.y += wgt * (p1.y + p2.y + p3.y);
    }
    float ia = 1.0f / area;
    c.x *= ia;
    c.y *= ia;
    h.centroid = c;
}

struct Mass {
    float mass, cx, cy, I;
};

// b2PolygonShape::ComputeMass
Mass hullMass(const Hull& h, float density) {
    P2 center{0.0f, 0.0f};
    float area = 0.0f, I = 0.0f;
    P2 s{0.0f, 0.0f};
    for (int i = 0; i < h.n; ++i) { s.x += h.v[i].x; s.y += h.v[i].y; }
    float invn = 1.0f / h.n;
    s.x *= invn;
    s.y *= invn;
    const float third = 1.0f / 3.0f;
    for (int i = 0; i < h.n; ++i) {
        P2 e1 = sub(h.v[i], s);
        P2 e2 = (i + 1 < h.n) ? sub(h.v[i + 1], s) : sub(h.v[0], s);
        float D = cross2(e1, e2);
        float tri = 0.5f * D;
        area += tri;
        float wgt = tri * third;
        center.x += wgt * (e1.x + e2.x);
        center.y += wgt * (e1.y + e2.y);
        float intx2 = e1.x * e1.x + e2.x * e1.x + e2.x * e2.x;
        float inty2 = e1.y * e1.y + e2.y * e1.y + e2.y * e2.y;
        I += (0.25f * third * D) * (intx2 + inty2);
    }
    Mass m;
    m.mass = density * area;
    float ia = 1.0f / area;
    center.x *= ia;
    center.y *= ia;
    m.cx = center.x + s.x;
    m.cy = center.y + s.y;
    m.I = density * I;
    m.I += m.mass * ((m.cx * m.cx + m.cy * m.cy) - (center.x * center.x + center.y * center.y));
    return m;
}

// boost::geometry::area of the (corrected) polygon; its surveyor strategy accumulates float
// coordinates in double (Box2DWorld.cpp:94-101)
double polygonArea(const std::vector<float>& p) {
    size_t n = p.size() / 2;
    double acc = 0.0;
    for (size_t i = 0; i < n; ++i) {
        size_t j = (i + 1 == n) ? 0 : i + 1;
        acc += ((double)p[2 * i] + (double)p[2 * j]) * ((double)p[2 * i + 1] - (double)p[2 * j + 1]);
    }
    acc /= 2.0;
    return acc < 0.0 ? -acc : acc;
}

struct BBody {
    int type = 0;  // 0 static, 2 dynamic
    float mass = 0, invMass = 0, I = 0, invI = 0, lcx = 0, lcy = 0;
    int fixStart = 0, fixCount = 0, object = -1;
    float px = 0, py = 0;  // creation position (ground only)
    std::string name;
};
struct BFix {
    int body;
    Hull hull;
    float friction, restitution, density;
    bool sensor;
};
struct BJoint {
    int type, bodyA, bodyB, flags;
    float lax, lay, lbx, lby, ref, lower, upper, maxForce, maxTorque;
    float axisX = 0.0f, axisY = 0.0f;  // prismatic joints: normalised local axis in bodyA
    int dof = -1, object = -1;
    std::string name;  // revolute joints: the YAML name; ground-friction joints have none
};
struct BLink {
    int body;
    int parentJoint = -1;  // model joint
    int parentPos = -1;
    int subtree = 1;
    std::vector<int> chain;
};

uint32_t fbits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

}  // namespace

void layoutState(ModelHeader& h, int maxContacts) {
    h.maxc = maxContacts;
    int s = 0;  // 16-byte chunks
    h.cBody = s; s += h.nb * SB_CHUNKS;
    h.cFix = s; s += h.nf;
    h.cJoint = s; s += h.nj * SJ_CHUNKS;
    h.cJointW = s; s += h.nj * WJ_CHUNKS;
    h.cCont = s; s += h.maxc * SC_CHUNKS;
    h.cContW = s; s += h.maxc * WC_CHUNKS;
    h.cScalar = s; s += SS_CHUNKS;
    h.cTarget = s; s += (h.ntarget + 3) / 4;
    h.stateChunks = s;
}

namespace {
// A non-finite number anywhere in the description (`.inf`, `.nan`, an overflowing literal) would end up in the mass data or
// the world bounds and turn every world of every batch into NaNs after one step: refused with the place it was found.
void requireFinite(float v, const std::string& where) {
    if (!std::isfinite(v)) throw std::runtime_error("non-finite number in " + where);
}
void requireFinite(const EnvironmentDescription& env) {
    requireFinite(env.scale, "scale");
    for (float v : env.world_bounds) requireFinite(v, "world_bounds");
    auto object = [](const ObjectDescription& d) {
        for (const LinkDescription& l : d.links) {
            const std::string where = "link '" + l.name + "' of '" + d.name + "'";
            requireFinite(l.mass, where); requireFinite(l.ground_friction, where); requireFinite(l.ground_torque_integral, where);
            requireFinite(l.contact_friction, where); requireFinite(l.restitution, where);
            // upstream asserts density >= 0 (b2Body::CreateFixture) and force / torque limits >= 0 (b2FrictionJoint); a negative
            // contact friction makes b2MixFriction's square root a NaN
            if (l.mass < 0.0f || l.ground_friction < 0.0f || l.ground_torque_integral < 0.0f || l.contact_friction < 0.0f ||
                l.restitution < 0.0f)
                throw std::runtime_error("negative mass / friction / restitution in " + where);
            for (const auto& poly : l.polygons) for (float v : poly) requireFinite(v, "geometry of " + where);
            for (const auto& ball : l.balls) for (float v : ball) requireFinite(v, "balls of " + where);
        }
        for (const JointDescription& j : d.joints) {
            const std::string where = "joint '" + j.name + "' of '" + d.name + "'";
            for (int k = 0; k < 2; ++k) {
                requireFinite(j.axis[k], where); requireFinite(j.position_limits[k], where);
                requireFinite(j.velocity_limits[k], where); requireFinite(j.acceleration_limits[k], where);
            }
            requireFinite(j.axis_orientation, where);
        }
        for (int k = 0; k < 2; ++k) {
            requireFinite(d.translational_velocity_limits[k], "base_actuation of '" + d.name + "'");
            requireFinite(d.translational_acceleration_limits[k], "base_actuation of '" + d.name + "'");
            requireFinite(d.rotational_velocity_limits[k], "base_actuation of '" + d.name + "'");
            requireFinite(d.rotational_acceleration_limits[k], "base_actuation of '" + d.name + "'");
        }
    };
    for (const ObjectDescription& d : env.robots) object(d);
    for (const ObjectDescription& d : env.objects) object(d);
    for (const auto& kv : env.states) {
        for (float v : kv.second.configuration) requireFinite(v, "state of '" + kv.first + "'");
        for (float v : kv.second.velocity) requireFinite(v, "state of '" + kv.first + "'");
    }
}
}  // namespace

void buildModel(const EnvironmentDescription& env, int maxContacts, HostModel& out) {
    requireFinite(env);
    const float scale = env.scale;
    if (!(scale > 0.0f)) throw std::runtime_error("scale must be positive");
    std::vector<BBody> bodies;
    std::vector<float> linkAabb;    // per body: Box2DLink::_local_aabb (min x, min y, max x, max y; user units)
    std::vector<float> linkParams;  // per body: ground_friction, ground_torque_integral, contact_friction, friction joint
    std::vector<BFix> fixtures;
    std::vector<BJoint> joints;

    // ---- ground (createGround, Box2DWorld.cpp:3644-3662)
    float wb[4] = {env.world_bounds[0], env.world_bounds[1], env.world_bounds[2], env.world_bounds[3]};
    float nrm = sqrtf(wb[0] * wb[0] + wb[1] * wb[1] + wb[2] * wb[2] + wb[3] * wb[3]);
    if (nrm == 0.0) {
        wb[0] = scale * -100.0f; wb[1] = scale * -100.0f; wb[2] = scale * 100.0f; wb[3] = scale * 100.0f;
    }
    {
        BBody g;
        g.type = 0;
        g.px = 0.5f * scale * (wb[0] + wb[2]);
        g.py = 0.5f * scale * (wb[1] + wb[3]);
        g.fixStart = 0;
        g.fixCount = 1;
        g.name = "GROUND_BODY";
        bodies.push_back(g);
        BFix f;
        f.body = 0;
        boxHull(f.hull, scale * (wb[2] - wb[0]), scale * (wb[3] - wb[1]));
        f.friction = 0.2f;
        f.restitution = 0.0f;
        f.density = 0.0f;
        f.sensor = true;
        fixtures.push_back(f);
    }

    struct BObject {
        const ObjectDescription* d;
        int firstBody, nLinks;
        int baseBody;
        float originX = 0, originY = 0, mass = 0;
        std::vector<int> jointList;      // model joints, YAML order
        std::vector<int> jointChildPos;  // pre-order position of the child link
        std::vector<BLink> links;        // pre-order
        std::vector<int> nameOrder;      // bodies in link-name order
        int nDofs = 0, dofOffset = 0;
        bool hasPrismatic = false;
    };
    std::vector<BObject> objects;
    std::vector<std::string> extraWarnings;
    std::vector<const ObjectDescription*> descs;
    for (auto& r : env.robots) descs.push_back(&r);
    for (auto& o : env.objects) descs.push_back(&o);
    {
        std::map<std::string, int> seen;
        for (auto* d : descs)
            if (seen[d->name]++) throw std::runtime_error("duplicate object name '" + d->name + "'");
    }

    int nq = 0;
    for (size_t oi = 0; oi < descs.size(); ++oi) {
        const ObjectDescription& d = *descs[oi];
        if (d.links.empty()) throw std::runtime_error("object '" + d.name + "' has no links");
        if ((int)d.links.size() > kMaxLinksPerObject) throw std::runtime_error("object '" + d.name + "' has too many links");
        BObject ob;
        ob.d = &d;
        ob.firstBody = (int)bodies.size();
        ob.nLinks = (int)d.links.size();
        std::map<std::string, int> linkBody;
        std::map<std::string, std::pair<float, float>> aabbMin, aabbMax;
        // ---- links (Box2DLink ctor :44-148)
        for (const LinkDescription& ld : d.links) {
            if (linkBody.count(ld.name)) throw std::runtime_error("duplicate link name '" + ld.name + "'");
            BBody b;
            b.type = d.is_static ? 0 : 2;
            b.object = (int)oi;
            b.fixStart = (int)fixtures.size();
            b.name = d.name + "/" + ld.name;
            float area = 0.0f;
            float mnx = std::numeric_limits<float>::max(), mny = mnx;
            float mxx = std::numeric_limits<float>::lowest(), mxy = mxx;
            std::vector<Hull> hulls;
            if (ld.polygons.empty()) throw std::runtime_error("link '" + ld.name + "' has no geometry");
            for (const std::vector<float>& poly : ld.polygons) {
                std::vector<P2> pts;
                for (size_t i = 0; i < poly.size() / 2; ++i) {
                    mnx = std::min(poly[2 * i], mnx);
                    mny = std::min(poly[2 * i + 1], mny);
                    mxx = std::max(poly[2 * i], mxx);
                    mxy = std::max(poly[2 * i + 1], mxy);
                    pts.push_back(P2{scale * poly[2 * i], scale * poly[2 * i + 1]});
                }
                area += scale * scale * polygonArea(poly);
                Hull h;
                convexHull(h, pts);
                hulls.push_back(h);
            }
            if (!(area > 0.0f)) throw std::runtime_error("link '" + ld.name + "' has zero area");
            float density = ld.mass / area;
            for (const Hull& h : hulls) {
                BFix f;
                f.body = (int)bodies.size();
                f.hull = h;
                f.friction = ld.contact_friction;
                f.restitution = ld.restitution;
                f.density = density;
                f.sensor = false;
                fixtures.push_back(f);
            }
            b.fixCount = (int)hulls.size();
            // b2Body::ResetMassData after the last CreateFixture (fixture list is newest first)
            if (b.type == 2) {
                float m = 0.0f, I = 0.0f, lx = 0.0f, ly = 0.0f;
                for (int k = (int)hulls.size() - 1; k >= 0; --k) {
                    if (density == 0.0f) continue;
                    Mass md = hullMass(hulls[k], density);
                    m += md.mass;
                    lx += md.mass * md.cx;
                    ly += md.mass * md.cy;
                    I += md.I;
                }
                if (m > 0.0f) {
                    b.invMass = 1.0f / m;
                    lx *= b.invMass;
                    ly *= b.invMass;
                } else {
                    m = 1.0f;
                    b.invMass = 1.0f;
                }
                if (I > 0.0f) {
                    I -= m * (lx * lx + ly * ly);
                    b.invI = 1.0f / I;
                } else {
                    I = 0.0f;
                    b.invI = 0.0f;
                }
                b.mass = m;
                b.I = I;
                b.lcx = lx;
                b.lcy = ly;
            }
            linkBody[ld.name] = (int)bodies.size();
            aabbMin[ld.name] = {mnx, mny};
            aabbMax[ld.name] = {mxx, mxy};
            linkAabb.resize(4 * bodies.size(), 0.0f);
            linkAabb.push_back(mnx); linkAabb.push_back(mny); linkAabb.push_back(mxx); linkAabb.push_back(mxy);
            ob.mass += b.mass * (b.type == 2 ? 1.0f : 0.0f);  // b2Body::GetMass() of a static body is 0
            // friction joint to the ground (:131-147)
            BJoint fj;
            fj.type = 0;
            fj.bodyA = 0;
            fj.bodyB = (int)bodies.size();
            fj.flags = 0;
            fj.lax = 0.0f; fj.lay = 0.0f;
            fj.lbx = b.lcx; fj.lby = b.lcy;
            fj.ref = 0.0f; fj.lower = 0.0f; fj.upper = 0.0f;
            float gravity = 9.81f * scale;
            float bodyMass = (b.type == 2) ? b.mass : 0.0f;
            fj.maxForce = ld.ground_friction * gravity * bodyMass;
            fj.maxTorque = ld.ground_friction * ld.ground_torque_integral * scale * scale;
            fj.object = (int)oi;
            linkParams.resize(4 * bodies.size(), 0.0f);  // rows of bodies without a link (the ground) stay zero / -1
            linkParams.push_back(ld.ground_friction);
            linkParams.push_back(ld.ground_torque_integral);
            linkParams.push_back(ld.contact_friction);
            linkParams.push_back((float)joints.size());
            joints.push_back(fj);
            bodies.push_back(b);
        }
        // Box2DObject::_mass accumulates link->getMass() in YAML order (:1303)
        {
            float m = 0.0f;
            for (int k = 0; k < ob.nLinks; ++k) {
                const BBody& b = bodies[ob.firstBody + k];
                m += (b.type == 2) ? b.mass : 0.0f;
            }
            ob.mass = m;
        }
        // ---- base link and origin (:1307-1312, 653-671)
        std::string base = d.base_link;
        if (base.empty()) {
            ObjectDescription tmp = d;
            std::string msg;
            if (!verifyKinematicStructure(tmp, msg)) throw std::runtime_error("object '" + d.name + "': " + msg);
            base = tmp.base_link;
        }
        if (!linkBody.count(base)) throw std::runtime_error("object '" + d.name + "': unknown base link '" + base + "'");
        ob.baseBody = linkBody[base];
        if (d.is_static) {
            float cx = 0.5f * (aabbMin[base].first + aabbMax[base].first);
            float cy = 0.5f * (aabbMin[base].second + aabbMax[base].second);
            ob.originX = scale * cx;
            ob.originY = scale * cy;
        } else {
            ob.originX = bodies[ob.baseBody].lcx;
            ob.originY = bodies[ob.baseBody].lcy;
        }
        // ---- joints (Box2DJoint ctor :872-948)
        int baseDofs = d.is_static ? 0 : 3;
        ob.dofOffset = nq;
        std::map<int, std::vector<std::pair<int, int>>> children;  // parent body -> (joint, child body), registration order
        std::map<int, int> parentJointOf;
        for (const JointDescription& jd : d.joints) {
            // Box2DJoint ctor (:888-898): "revolute" / "prismatic"; anything else is reported and treated as revolute
            const bool prismatic = jd.joint_type == "prismatic";
            if (!prismatic && jd.joint_type != "revolute")
                extraWarnings.push_back("Unknown joint type " + jd.joint_type + ". Using revolute instead.");
            if (!linkBody.count(jd.link_a) || !linkBody.count(jd.link_b))
                throw std::runtime_error("joint '" + jd.name + "' names an unknown link");
            BJoint j;
            j.type = prismatic ? 2 : 1;
            j.bodyA = linkBody[jd.link_a];
            j.bodyB = linkBody[jd.link_b];
            j.lax = scale * jd.axis[0];
            j.lay = scale * jd.axis[1];
            j.lbx = 0.0f;
            j.lby = 0.0f;
            j.ref = jd.axis_orientation;
            j.lower = jd.position_limits[0];
            j.upper = jd.position_limits[1];
            float n2 = sqrtf(jd.position_limits[0] * jd.position_limits[0] + jd.position_limits[1] * jd.position_limits[1]);
            j.flags = (n2 > 0.0 ? 1 : 0) | (jd.actuated ? 2 : 0);
            j.maxForce = 0.0f;
            j.maxTorque = 0.0f;
            if (prismatic) {
                // b2PrismaticJointDef as the reference fills it (:920-940): localAnchorA is NOT scaled, the translation limits
                // are; the axis is the unit vector at axis_orientation (libm cos / sin of a description constant), normalised
                // by b2PrismaticJoint's constructor; maxMotorForce = min |acceleration limits| (kept in maxForce)
                j.lax = jd.axis[0];
                j.lay = jd.axis[1];
                j.lower = scale * jd.position_limits[0];
                j.upper = scale * jd.position_limits[1];
                float ax = std::cos(jd.axis_orientation), ay = std::sin(jd.axis_orientation);
                float len = sqrtf(ax * ax + ay * ay);
                if (!(len < FLT_EPSILON)) {
                    float inv = 1.0f / len;
                    ax *= inv;
                    ay *= inv;
                }
                j.axisX = ax;
                j.axisY = ay;
                j.maxForce = std::min(std::abs(jd.acceleration_limits[0]), std::abs(jd.acceleration_limits[1]));
                ob.hasPrismatic = true;
                if (d.is_robot)
                    throw std::runtime_error("unsupported: robot '" + d.name + "' has the prismatic joint '" + jd.name + "': the reference "
                                             "cannot control it (Box2DJoint::setControlTorque throws, Box2DWorld.cpp:1204-1208)");
            }
            j.dof = nq + baseDofs + (int)ob.jointList.size();
            j.object = (int)oi;
            j.name = jd.name;
            children[j.bodyA].push_back({(int)joints.size(), j.bodyB});
            parentJointOf[j.bodyB] = (int)joints.size();
            ob.jointList.push_back((int)joints.size());
            joints.push_back(j);
        }
        ob.nDofs = baseDofs + (int)ob.jointList.size();
        if (ob.nDofs > kMaxDofsPerObject) throw std::runtime_error("object '" + d.name + "' has too many DOFs");
        nq += ob.nDofs;
        // ---- pre-order link list with ancestor chains
        {
            struct Item { int body, parentJoint, parentPos; std::vector<int> chain; };
            std::vector<Item> stack;
            stack.push_back(Item{ob.baseBody, -1, -1, {}});
            while (!stack.empty()) {
                Item it = stack.back();
                stack.pop_back();
                BLink l;
                l.body = it.body;
                l.parentJoint = it.parentJoint;
                l.parentPos = it.parentPos;
                l.chain = it.chain;
                int pos = (int)ob.links.size();
                ob.links.push_back(l);
                auto& ch = children[it.body];
                for (int k = (int)ch.size() - 1; k >= 0; --k) {  // reversed push -> registration order when popped
                    std::vector<int> c2 = it.chain;
                    c2.push_back(ch[k].first);
                    stack.push_back(Item{ch[k].second, ch[k].first, pos, c2});
                }
            }
            if ((int)ob.links.size() != ob.nLinks)
                throw std::runtime_error("object '" + d.name + "': links are not one kinematic tree");
            for (int p = ob.nLinks - 1; p >= 0; --p)
                if (ob.links[p].parentPos >= 0) ob.links[ob.links[p].parentPos].subtree += ob.links[p].subtree;
            for (int jm : ob.jointList) {
                int child = joints[jm].bodyB;
                int pos = -1;
                for (int p = 0; p < ob.nLinks; ++p)
                    if (ob.links[p].body == child) pos = p;
                ob.jointChildPos.push_back(pos);
            }
        }
        for (auto& kv : linkBody) ob.nameOrder.push_back(kv.second);  // std::map iterates in name order
        objects.push_back(ob);
    }

    int nb = (int)bodies.size(), nf = (int)fixtures.size(), nj = (int)joints.size(), no = (int)objects.size();
    if (nb > kMaxBodies) throw std::runtime_error("too many bodies (max " + std::to_string(kMaxBodies) + ")");
    if (nf > kMaxFixtures) throw std::runtime_error("too many fixtures (max " + std::to_string(kMaxFixtures) + ")");
    if (nj > kMaxJoints) throw std::runtime_error("too many joints (max " + std::to_string(kMaxJoints) + ")");

    // ---- joint edges per body, newest joint first; collide masks
    std::vector<std::vector<std::pair<int, int>>> jedges(nb);
    std::vector<uint64_t> collide(nb, 0);
    for (int j = nj - 1; j >= 0; --j) {
        jedges[joints[j].bodyA].push_back({j, joints[j].bodyB});
        jedges[joints[j].bodyB].push_back({j, joints[j].bodyA});
    }
    for (int a = 0; a < nb; ++a)
        for (int b = 0; b < nb; ++b) {
            if (a == b) continue;
            if (bodies[a].type != 2 && bodies[b].type != 2) continue;
            bool jointed = false;
            for (auto& e : jedges[a])
                if (e.second == b) jointed = true;  // every joint the reference creates has collideConnected == false
            if (!jointed) collide[a] |= (1ull << b);
        }
    std::vector<uint64_t> pairMask(nf, 0);
    int nPairs = 0;
    for (int f = 0; f < nf; ++f)
        for (int g = 0; g < nf; ++g) {
            if (f == g) continue;
            int ba = fixtures[f].body, bb = fixtures[g].body;
            if (ba == bb) continue;
            if (!(collide[ba] >> bb & 1ull)) continue;
            pairMask[f] |= (1ull << g);
            if (f < g) ++nPairs;
        }

    // ---- robots in name order (std::map<std::string, Box2DRobotPtr> _robots, Box2DWorld.h)
    std::vector<int> robotOrder;
    {
        std::map<std::string, int> byName;
        for (int oi = 0; oi < no; ++oi)
            if (objects[oi].d->is_robot) byName[objects[oi].d->name] = oi;
        for (auto& kv : byName) robotOrder.push_back(kv.second);
    }
    std::vector<int> targetOffset(no, -1);
    int ntarget = 0;
    for (int oi = 0; oi < no; ++oi)
        if (objects[oi].d->is_robot) {
            targetOffset[oi] = ntarget;
            ntarget += objects[oi].nDofs + 1;  // the controller targets of the robot's DOFs + the position controller's gain
        }

    // ---- flatten
    ModelHeader& h = out.h;
    memset(&h, 0, sizeof(h));
    if (robotOrder.size() > 8) throw std::runtime_error("too many robots (max 8: controller mode bits per world)");
    h.nb = nb; h.nf = nf; h.nj = nj; h.no = no; h.nrobots = (int)robotOrder.size(); h.nq = nq; h.ntarget = ntarget;
    h.scale = scale;
    h.invScale = 1.0f / scale;
    for (const auto& ob : objects) h.nPrismatic += ob.hasPrismatic ? 1 : 0;
    std::vector<uint32_t>& B = out.blob;
    B.clear();
    auto allocWords = [&](int n) { int o = (int)B.size(); B.resize(B.size() + n, 0u); return o; };
    auto align4 = [&]() { while (B.size() % 4) B.push_back(0u); };  // records read with LDS.128 start on 16 bytes

    h.oJEdge = allocWords(0);
    std::vector<int> jeStart(nb);
    for (int b = 0; b < nb; ++b) {
        jeStart[b] = (int)B.size() - h.oJEdge;
        for (auto& e : jedges[b]) B.push_back((uint32_t)(e.first | (e.second << 16)));
    }
    align4();
    h.oBody = allocWords(nb * MB_STRIDE);
    for (int b = 0; b < nb; ++b) {
        uint32_t* w = &B[h.oBody + b * MB_STRIDE];
        const BBody& bd = bodies[b];
        w[MB_INVMASS] = fbits(bd.invMass); w[MB_INVI] = fbits(bd.invI);
        w[MB_MASS] = fbits(bd.type == 2 ? bd.mass : 0.0f); w[MB_I] = fbits(bd.I);
        w[MB_LCX] = fbits(bd.lcx); w[MB_LCY] = fbits(bd.lcy);
        w[MB_TYPE] = bd.type; w[MB_JE_START] = jeStart[b]; w[MB_JE_COUNT] = (uint32_t)jedges[b].size();
        w[MB_FIX_START] = bd.fixStart; w[MB_FIX_COUNT] = bd.fixCount; w[MB_OBJECT] = (uint32_t)bd.object;
        w[MB_COLLIDE_LO] = (uint32_t)(collide[b] & 0xffffffffull); w[MB_COLLIDE_HI] = (uint32_t)(collide[b] >> 32);
        w[MB_PX] = fbits(bd.px); w[MB_PY] = fbits(bd.py);
        w[MB_SOFF] = (uint32_t)(b * SB_CHUNKS * kChunkBytes);  // layoutState puts the bodies first (cBody == 0)
        w[MB_FORCED] = 0u;
        for (const BObject& ob : objects)
            if (ob.d->is_robot && ob.baseBody == b) w[MB_FORCED] = 1u;
    }
    align4();
    h.oFix = allocWords(nf * MF_STRIDE);
    for (int f = 0; f < nf; ++f) {
        uint32_t* w = &B[h.oFix + f * MF_STRIDE];
        const BFix& fx = fixtures[f];
        w[MF_BODY] = fx.body; w[MF_COUNT] = fx.hull.n;
        w[MF_FRICTION] = fbits(fx.friction); w[MF_RESTITUTION] = fbits(fx.restitution);
        w[MF_PAIR_LO] = (uint32_t)(pairMask[f] & 0xffffffffull); w[MF_PAIR_HI] = (uint32_t)(pairMask[f] >> 32);
        // fat AABB at creation: b2PolygonShape::ComputeAABB with the body at its creation pose (angle 0)
        float px = bodies[fx.body].px, py = bodies[fx.body].py;
        float lx = 0, ly = 0, ux = 0, uy = 0;
        for (int i = 0; i < fx.hull.n; ++i) {
            float x = (1.0f * fx.hull.v[i].x - 0.0f * fx.hull.v[i].y) + px;
            float y = (0.0f * fx.hull.v[i].x + 1.0f * fx.hull.v[i].y) + py;
            if (i == 0) { lx = ux = x; ly = uy = y; }
            else {
                lx = lx < x ? lx : x; ly = ly < y ? ly : y;
                ux = ux > x ? ux : x; uy = uy > y ? uy : y;
            }
        }
        lx -= kPolygonRadius; ly -= kPolygonRadius; ux += kPolygonRadius; uy += kPolygonRadius;
        w[MF_FAT0 + 0] = fbits(lx - kAabbExtension); w[MF_FAT0 + 1] = fbits(ly - kAabbExtension);
        w[MF_FAT0 + 2] = fbits(ux + kAabbExtension); w[MF_FAT0 + 3] = fbits(uy + kAabbExtension);
        for (int i = 0; i < kMaxPolyVerts; ++i) {
            P2 v = i < fx.hull.n ? fx.hull.v[i] : P2{0.0f, 0.0f};
            P2 n = i < fx.hull.n ? fx.hull.nrm[i] : P2{0.0f, 0.0f};
            w[MF_VERTS + 2 * i] = fbits(v.x); w[MF_VERTS + 2 * i + 1] = fbits(v.y);
            w[MF_NORMALS + 2 * i] = fbits(n.x); w[MF_NORMALS + 2 * i + 1] = fbits(n.y);
        }
    }
    align4();
    h.oJoint = allocWords(nj * MJ_STRIDE);
    for (int j = 0; j < nj; ++j) {
        uint32_t* w = &B[h.oJoint + j * MJ_STRIDE];
        const BJoint& jt = joints[j];
        w[MJ_TYPE] = jt.type; w[MJ_BODYA] = jt.bodyA; w[MJ_BODYB] = jt.bodyB; w[MJ_FLAGS] = jt.flags;
        w[MJ_LAX] = fbits(jt.lax); w[MJ_LAY] = fbits(jt.lay); w[MJ_LBX] = fbits(jt.lbx); w[MJ_LBY] = fbits(jt.lby);
        w[MJ_REF] = fbits(jt.ref); w[MJ_LOWER] = fbits(jt.lower); w[MJ_UPPER] = fbits(jt.upper);
        w[MJ_MAXFORCE] = fbits(jt.maxForce); w[MJ_MAXTORQUE] = fbits(jt.maxTorque);
        w[MJ_DOF] = (uint32_t)jt.dof; w[MJ_OBJECT] = (uint32_t)jt.object;
        w[MJ_AXISX] = fbits(jt.axisX); w[MJ_AXISY] = fbits(jt.axisY);
        const BBody& A = bodies[jt.bodyA];
        const BBody& Bd = bodies[jt.bodyB];
        if (A.type == 0) w[MJ_FLAGS] |= 4u;
        if (Bd.type == 0) w[MJ_FLAGS] |= 8u;
        // state offsets: bodies, fixtures, joint state, joint scratch come first in layoutState and do not depend on maxc
        const int cJoint = nb * SB_CHUNKS + nf, cJointW = cJoint + nj * SJ_CHUNKS;
        w[MJ_SOFF] = (uint32_t)((cJoint + j * SJ_CHUNKS) * kChunkBytes);
        w[MJ_WOFF] = (uint32_t)((cJointW + j * WJ_CHUNKS) * kChunkBytes);
        w[MJ_BAOFF] = (uint32_t)(jt.bodyA * SB_CHUNKS * kChunkBytes);
        w[MJ_BBOFF] = (uint32_t)(jt.bodyB * SB_CHUNKS * kChunkBytes);
        w[MJ_RA0X] = fbits(jt.lax - A.lcx); w[MJ_RA0Y] = fbits(jt.lay - A.lcy);
        w[MJ_RB0X] = fbits(jt.lbx - Bd.lcx); w[MJ_RB0Y] = fbits(jt.lby - Bd.lcy);
        w[MJ_MA] = fbits(A.invMass); w[MJ_IA] = fbits(A.invI); w[MJ_MB] = fbits(Bd.invMass); w[MJ_IB] = fbits(Bd.invI);
        {   // b2FrictionJoint::m_angularMass / b2RevoluteJoint::m_motorMass: constant for the joint
            float m = A.invI + Bd.invI;
            if (m > 0.0f) m = 1.0f / m;
            w[MJ_ROTMASS] = fbits(m);
        }
        // the compact friction-joint layout (model.h) relies on what Box2DLink's ctor always builds: ground -> link
        if (jt.type == 0 && A.type != 0) throw std::runtime_error("friction joints must hang off the static ground body");
    }
    // links / chains / object joints / name order / init
    h.oLink = allocWords(0);
    std::vector<int> linkStart(no);
    std::vector<std::vector<int>> chainStarts(no);
    {
        int total = 0;
        for (int oi = 0; oi < no; ++oi) { linkStart[oi] = total; total += objects[oi].nLinks; }
        allocWords(total * ML_STRIDE);
    }
    h.oChain = allocWords(0);
    for (int oi = 0; oi < no; ++oi)
        for (int p = 0; p < objects[oi].nLinks; ++p) {
            const BLink& l = objects[oi].links[p];
            uint32_t* w = &B[h.oLink + (linkStart[oi] + p) * ML_STRIDE];
            w[ML_BODY] = l.body; w[ML_PARENT_JOINT] = (uint32_t)l.parentJoint; w[ML_PARENT_POS] = (uint32_t)l.parentPos;
            w[ML_SUBTREE] = l.subtree; w[ML_CHAIN_START] = (int)B.size() - h.oChain; w[ML_CHAIN_LEN] = (uint32_t)l.chain.size();
            for (int c : l.chain) B.push_back((uint32_t)c);
            // (B may reallocate: w is not used after push_back)
        }
    h.oObjJoint = allocWords(0);
    std::vector<int> ojStart(no);
    for (int oi = 0; oi < no; ++oi) {
        ojStart[oi] = (int)B.size() - h.oObjJoint;
        for (size_t k = 0; k < objects[oi].jointList.size(); ++k)
            B.push_back((uint32_t)(objects[oi].jointList[k] | (objects[oi].jointChildPos[k] << 16)));
    }
    h.oName = allocWords(0);
    std::vector<int> nameStart(no);
    for (int oi = 0; oi < no; ++oi) {
        nameStart[oi] = (int)B.size() - h.oName;
        for (int b : objects[oi].nameOrder) B.push_back((uint32_t)b);
    }
    h.oRobotOrder = allocWords(0);
    for (int r : robotOrder) B.push_back((uint32_t)r);
    h.oInit = allocWords(0);
    std::vector<int> initStart(no);
    for (int oi = 0; oi < no; ++oi) {
        initStart[oi] = (int)B.size() - h.oInit;
        auto it = env.states.find(objects[oi].d->name);
        const BObject& ob = objects[oi];
        if (it == env.states.end()) {
            B.push_back(0u);
            B.push_back(0u);
            continue;
        }
        const StateDescription& sd = it->second;
        if (ob.hasPrismatic)
            throw std::runtime_error("unsupported: object '" + ob.d->name + "' has a prismatic joint and a `states` entry: the reference "
                                     "throws while applying it (Box2DLink::updateBodyVelocities -> getGlobalAxisPosition, "
                                     "Box2DWorld.cpp:789, 1282-1285)");
        int expect = ob.d->is_static ? ob.nDofs + 3 : ob.nDofs;
        if ((int)sd.configuration.size() != expect || (int)sd.velocity.size() != expect)
            throw std::runtime_error("state of '" + ob.d->name + "' needs " + std::to_string(expect) + " values");
        B.push_back(1u);
        B.push_back((uint32_t)expect);
        for (float v : sd.configuration) B.push_back(fbits(v));
        for (float v : sd.velocity) B.push_back(fbits(v));
    }
    h.oDof = allocWords(nq * 6);
    out.dofLimits.assign(nq * 6, 0.0f);
    {
        const float lo = std::numeric_limits<float>::lowest(), hi = std::numeric_limits<float>::max();
        for (int oi = 0; oi < no; ++oi) {
            const BObject& ob = objects[oi];
            const ObjectDescription& d = *ob.d;
            int baseDofs = d.is_static ? 0 : 3;
            for (int k = 0; k < ob.nDofs; ++k) {
                float lim[6] = {lo, hi, lo, hi, lo, hi};
                if (k < baseDofs) {
                    if (k == 2) { lim[0] = (float)-M_PI; lim[1] = (float)M_PI; }
                    if (d.is_robot) {  // Box2DRobot ctor :2056-2066
                        const float* vl = k < 2 ? d.translational_velocity_limits : d.rotational_velocity_limits;
                        const float* al = k < 2 ? d.translational_acceleration_limits : d.rotational_acceleration_limits;
                        lim[2] = vl[0]; lim[3] = vl[1]; lim[4] = al[0]; lim[5] = al[1];
                    }
                } else {
                    const JointDescription& jd = d.joints[k - baseDofs];
                    lim[0] = jd.position_limits[0]; lim[1] = jd.position_limits[1];
                    lim[2] = jd.velocity_limits[0]; lim[3] = jd.velocity_limits[1];
                    lim[4] = jd.acceleration_limits[0]; lim[5] = jd.acceleration_limits[1];
                }
                for (int c = 0; c < 6; ++c) {
                    B[h.oDof + (ob.dofOffset + k) * 6 + c] = fbits(lim[c]);
                    out.dofLimits[(ob.dofOffset + k) * 6 + c] = lim[c];
                }
            }
        }
    }
    h.oObj = allocWords(no * MO_STRIDE);
    for (int oi = 0; oi < no; ++oi) {
        uint32_t* w = &B[h.oObj + oi * MO_STRIDE];
        const BObject& ob = objects[oi];
        w[MO_STATIC] = ob.d->is_static; w[MO_ROBOT] = ob.d->is_robot; w[MO_BASE_BODY] = ob.baseBody;
        w[MO_NDOFS] = ob.nDofs; w[MO_DOF_OFFSET] = ob.dofOffset;
        w[MO_LINK_START] = linkStart[oi]; w[MO_LINK_COUNT] = ob.nLinks;
        w[MO_JOINT_START] = ojStart[oi]; w[MO_JOINT_COUNT] = (uint32_t)ob.jointList.size();
        w[MO_ORIGIN_X] = fbits(ob.originX); w[MO_ORIGIN_Y] = fbits(ob.originY); w[MO_MASS] = fbits(ob.mass);
        w[MO_NAME_START] = nameStart[oi]; w[MO_TARGET_OFFSET] = (uint32_t)targetOffset[oi];
        w[MO_INIT_START] = initStart[oi];
    }
    h.words = (int)B.size();

    if (maxContacts <= 0) maxContacts = std::min(nPairs, std::max(32, 2 * nf));
    if (maxContacts < 1) maxContacts = 1;
    if (maxContacts > kMaxContacts) maxContacts = kMaxContacts;
    layoutState(h, maxContacts);

    // ---- introspection
    out.objects.clear();
    out.balls.clear();
    for (int oi = 0; oi < no; ++oi) {
        HostObjectInfo info;
        info.name = objects[oi].d->name;
        info.isRobot = objects[oi].d->is_robot;
        info.isStatic = objects[oi].d->is_static;
        info.nDofs = objects[oi].nDofs;
        info.dofOffset = objects[oi].dofOffset;
        info.firstBody = objects[oi].firstBody;
        info.nLinks = objects[oi].nLinks;
        info.nJoints = (int)objects[oi].jointList.size();
        info.hasPrismatic = objects[oi].hasPrismatic;
        info.ballStart = (int)out.balls.size() / 8;
        for (int b : objects[oi].nameOrder) {
            const LinkDescription& ld = objects[oi].d->links[b - objects[oi].firstBody];
            const bool base = b == objects[oi].baseBody;  // only the base link has a _local_origin_offset (:53, 653-671)
            for (auto& ball : ld.balls) {
                const float row[8] = {(float)b, base ? objects[oi].originX : 0.0f, base ? objects[oi].originY : 0.0f,
                                      ball[0], ball[1], ball[2], ball[3], 0.0f};
                out.balls.insert(out.balls.end(), row, row + 8);
            }
        }
        info.nBalls = (int)out.balls.size() / 8 - info.ballStart;
        out.objects.push_back(info);
    }
    out.bodyNames.clear();
    out.bodyModel.assign(nb * 8, 0.0f);
    for (int b = 0; b < nb; ++b) {
        out.bodyNames.push_back(bodies[b].name);
        float* o = &out.bodyModel[b * 8];
        o[0] = bodies[b].type == 2 ? bodies[b].mass : 0.0f; o[1] = bodies[b].invMass; o[2] = bodies[b].I; o[3] = bodies[b].invI;
        o[4] = bodies[b].lcx; o[5] = bodies[b].lcy; o[6] = (float)bodies[b].type; o[7] = (float)bodies[b].fixCount;
    }
    out.fixInts.assign(nf * 3, 0);
    out.fixFloats.assign(nf * 38, 0.0f);
    for (int f = 0; f < nf; ++f) {
        out.fixInts[f * 3] = fixtures[f].body; out.fixInts[f * 3 + 1] = fixtures[f].hull.n; out.fixInts[f * 3 + 2] = fixtures[f].sensor;
        float* o = &out.fixFloats[f * 38];
        for (int i = 0; i < fixtures[f].hull.n; ++i) {
            o[2 * i] = fixtures[f].hull.v[i].x; o[2 * i + 1] = fixtures[f].hull.v[i].y;
            o[16 + 2 * i] = fixtures[f].hull.nrm[i].x; o[16 + 2 * i + 1] = fixtures[f].hull.nrm[i].y;
        }
        o[32] = fixtures[f].hull.centroid.x; o[33] = fixtures[f].hull.centroid.y;
        o[34] = fixtures[f].friction; o[35] = fixtures[f].restitution; o[36] = fixtures[f].density; o[37] = kPolygonRadius;
    }
    out.jointInts.assign(nj * 6, 0);
    out.jointFloats.assign(nj * 15, 0.0f);
    for (int j = 0; j < nj; ++j) {
        int32_t* io = &out.jointInts[j * 6];
        float* fo = &out.jointFloats[j * 15];
        io[0] = joints[j].type; io[1] = joints[j].bodyA; io[2] = joints[j].bodyB; io[3] = 0;
        io[4] = (joints[j].flags >> 1) & 1; io[5] = joints[j].flags & 1;
        fo[6] = joints[j].maxForce; fo[7] = joints[j].maxTorque;
        fo[8] = joints[j].lax; fo[9] = joints[j].lay; fo[10] = joints[j].lbx; fo[11] = joints[j].lby;
        fo[12] = joints[j].ref; fo[13] = joints[j].lower; fo[14] = joints[j].upper;
    }
    out.warnings = env.warnings;
    out.warnings.insert(out.warnings.end(), extraWarnings.begin(), extraWarnings.end());
    out.jointNames.clear();
    for (int j = 0; j < nj; ++j) out.jointNames.push_back(joints[j].name);
    linkAabb.resize(4 * nb, 0.0f);
    out.linkAabb = linkAabb;
    linkParams.resize(4 * nb, 0.0f);
    out.linkParams = linkParams;
    for (int b = 0; b < nb; ++b) {
        bool isLink = false;
        for (int j = 0; j < nj; ++j) isLink = isLink || (joints[j].type == 0 && joints[j].bodyB == b);
        if (!isLink) out.linkParams[4 * b + 3] = -1.0f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Box2DLink's dynamics parameters on a built model (Box2DWorld.cpp:471-566).  Every quantity the kernels read that
// depends on them is re-derived in place: the body record (mass data), the records of the joints that touch the body
// (inverse masses, angular / motor mass) and the link's ground-friction joint (maxForce / maxTorque).
namespace {
float wordF(const HostModel& m, int word) {
    float f;
    std::memcpy(&f, &m.blob[word], 4);
    return f;
}
void setWordF(HostModel& m, int word, float f) { std::memcpy(&m.blob[word], &f, 4); }

void refreshJointsOfBody(HostModel& m, int body) {
    const ModelHeader& h = m.h;
    for (int j = 0; j < h.nj; ++j) {
        const int jr = h.oJoint + j * MJ_STRIDE;
        const int a = (int)m.blob[jr + MJ_BODYA], b = (int)m.blob[jr + MJ_BODYB];
        if (a != body && b != body) continue;
        const float mA = wordF(m, h.oBody + a * MB_STRIDE + MB_INVMASS), iA = wordF(m, h.oBody + a * MB_STRIDE + MB_INVI);
        const float mB = wordF(m, h.oBody + b * MB_STRIDE + MB_INVMASS), iB = wordF(m, h.oBody + b * MB_STRIDE + MB_INVI);
        setWordF(m, jr + MJ_MA, mA); setWordF(m, jr + MJ_IA, iA); setWordF(m, jr + MJ_MB, mB); setWordF(m, jr + MJ_IB, iB);
        float rot = iA + iB;
        if (rot > 0.0f) rot = 1.0f / rot;
        setWordF(m, jr + MJ_ROTMASS, rot);
    }
}

// Box2DLink::updateFrictionJoint (:523-531): mu * mass * gravity (the ctor multiplies in the order mu * gravity * mass, :141)
void updateFrictionJoint(HostModel& m, int body) {
    const ModelHeader& h = m.h;
    const int j = (int)m.linkParams[4 * body + 3];
    const float mu = m.linkParams[4 * body], integral = m.linkParams[4 * body + 1];
    const float gravity = 9.81f * h.scale;
    const float mass = wordF(m, h.oBody + body * MB_STRIDE + MB_MASS);  // b2Body::GetMass(): 0 for a static body
    const float maxForce = mu * mass * gravity;
    const float maxTorque = mu * integral * h.scale * h.scale;
    setWordF(m, h.oJoint + j * MJ_STRIDE + MJ_MAXFORCE, maxForce);
    setWordF(m, h.oJoint + j * MJ_STRIDE + MJ_MAXTORQUE, maxTorque);
    m.jointFloats[j * 15 + 6] = maxForce;
    m.jointFloats[j * 15 + 7] = maxTorque;
}
}  // namespace

// Box2DObject::getLocalAABB (Box2DWorld.cpp:1928-1931): the link boxes merged as they are, each in its own link frame
// (:1304), shifted by the base link's local centre of mass as it was before setOriginToCenterOfMass (:1308-1312; a static
// body has none, b2Body::GetLocalCenter() is zero).  BoundingBox::merge belongs to the absent sim_env package: assumed to
// be the component-wise min / max starting from an empty box.
void objectLocalAabb(const HostModel& m, int object, float out[4]) {
    const HostObjectInfo& o = m.objects[object];
    out[0] = out[1] = std::numeric_limits<float>::max();
    out[2] = out[3] = std::numeric_limits<float>::lowest();
    for (int b = o.firstBody; b < o.firstBody + o.nLinks; ++b) {
        out[0] = std::min(out[0], m.linkAabb[4 * b]); out[1] = std::min(out[1], m.linkAabb[4 * b + 1]);
        out[2] = std::max(out[2], m.linkAabb[4 * b + 2]); out[3] = std::max(out[3], m.linkAabb[4 * b + 3]);
    }
    const int base = (int)m.blob[m.h.oObj + object * MO_STRIDE + MO_BASE_BODY];
    const bool dynamic = m.blob[m.h.oBody + base * MB_STRIDE + MB_TYPE] == 2u;
    const float cx = dynamic ? m.h.invScale * wordF(m, m.h.oBody + base * MB_STRIDE + MB_LCX) : 0.0f;
    const float cy = dynamic ? m.h.invScale * wordF(m, m.h.oBody + base * MB_STRIDE + MB_LCY) : 0.0f;
    out[0] -= cx; out[1] -= cy; out[2] -= cx; out[3] -= cy;
}

bool isLinkBody(const HostModel& m, int body) {
    return body >= 0 && body < m.h.nb && 4 * (size_t)body + 3 < m.linkParams.size() && m.linkParams[4 * body + 3] >= 0.0f;
}

bool setLinkProperty(HostModel& m, int body, int prop, float value) {
    const ModelHeader& h = m.h;
    const int br = h.oBody + body * MB_STRIDE;
    switch (prop) {
        case LP_MASS: {  // Box2DLink::setMass (:478-489) = b2Body::GetMassData, scale mass and I, b2Body::SetMassData
            const bool dynamic = m.blob[br + MB_TYPE] == 2u;
            if (dynamic) {
                const float lcx = wordF(m, br + MB_LCX), lcy = wordF(m, br + MB_LCY);
                const float oldMass = wordF(m, br + MB_MASS);
                float dataI = wordF(m, br + MB_I) + oldMass * (lcx * lcx + lcy * lcy);
                const float ratio = value / oldMass;
                dataI = dataI * ratio;
                float mass = value;
                if (mass <= 0.0f) mass = 1.0f;
                const float invMass = 1.0f / mass;
                float I = 0.0f, invI = 0.0f;
                if (dataI > 0.0f) {
                    I = dataI - mass * (lcx * lcx + lcy * lcy);
                    invI = 1.0f / I;
                }
                setWordF(m, br + MB_MASS, mass); setWordF(m, br + MB_INVMASS, invMass);
                setWordF(m, br + MB_I, I); setWordF(m, br + MB_INVI, invI);
                float* o = &m.bodyModel[body * 8];
                o[0] = mass; o[1] = invMass; o[2] = I; o[3] = invI;
                refreshJointsOfBody(m, body);
            }
            updateFrictionJoint(m, body);
            return dynamic;  // the caller re-derives the sweep centre of dynamic bodies (b2Body::SetMassData's tail)
        }
        case LP_GROUND_FRICTION:
            m.linkParams[4 * body] = value;
            updateFrictionJoint(m, body);
            return false;
        case LP_GROUND_TORQUE_INTEGRAL:
            m.linkParams[4 * body + 1] = value;
            updateFrictionJoint(m, body);
            return false;
        case LP_CONTACT_FRICTION: {  // Box2DLink::setContactFriction (:538-547): b2Fixture::SetFriction on every fixture
            m.linkParams[4 * body + 2] = value;
            const int f0 = (int)m.blob[br + MB_FIX_START], nfix = (int)m.blob[br + MB_FIX_COUNT];
            for (int f = f0; f < f0 + nfix; ++f) {
                setWordF(m, h.oFix + f * MF_STRIDE + MF_FRICTION, value);
                m.fixFloats[f * 38 + 34] = value;
            }
            return false;
        }
        default:
            throw std::runtime_error("link property is read-only or unknown");
    }
}

float getLinkProperty(const HostModel& m, int body, int prop) {
    const ModelHeader& h = m.h;
    const int br = h.oBody + body * MB_STRIDE;
    const float mass = wordF(m, br + MB_MASS);
    const float lcx = wordF(m, br + MB_LCX), lcy = wordF(m, br + MB_LCY);
    const float inertia = wordF(m, br + MB_I) + mass * (lcx * lcx + lcy * lcy);  // b2Body::GetInertia
    switch (prop) {
        case LP_MASS: return mass;
        case LP_GROUND_FRICTION: return m.linkParams[4 * body];
        case LP_GROUND_TORQUE_INTEGRAL: return m.linkParams[4 * body + 1];
        case LP_CONTACT_FRICTION: return m.linkParams[4 * body + 2];
        case LP_INERTIA: return h.invScale * h.invScale * inertia;  // Box2DLink::getInertia (:549-555)
        case LP_COM_INERTIA: {                                        // Box2DLink::getCOMInertia (:557-566)
            const float comDist = std::sqrt(lcx * lcx + lcy * lcy);
            const float comInertia = inertia - mass * comDist * comDist;
            return h.invScale * h.invScale * comInertia;
        }
        case LP_LIMIT_FORCE: return m.linkParams[4 * body] * mass * 9.81f;                 // :507-510
        case LP_LIMIT_TORQUE: return m.linkParams[4 * body] * m.linkParams[4 * body + 1];  // :502-505
        default: throw std::runtime_error("unknown link property");
    }
}

}  // namespace b2g

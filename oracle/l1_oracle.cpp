// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into, imported by or
// executed from the product path (prototype2_b200/, include/).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load the library this file builds into.
//
// L1 = CPU RESTATEMENT of the reference's character-collision hot path with
// widened (uint32) collider indices, std containers, parametrised sub-step
// count and two arrangements:
//
//   LITERAL    one unified dynamic AABB tree holding mesh leaves AND capsule
//              leaves, characters processed sequentially, capsule-capsule in
//              tree order -- the reference exactly as written.  Pinned: the
//              tests require L1-LITERAL to be bit-identical to L0 (the
//              reference's own sources compiled verbatim, oracle/l0_driver.cpp)
//              in states AND traces on every scene L0 can hold.
//   CANONICAL  static-only tree (built by the same insertion algorithm, then
//              frozen), capsule neighbours in ascending index order from the
//              stored fat AABBs, dynamic pass OFF or JACOBI.  This is the
//              arrangement the GPU path implements at scale.
//   BY_ID      CANONICAL, except that the mesh candidates of a query -- the
//              same SET, which does not depend on the tree's shape -- are
//              taken in ascending collider id (registration order) instead
//              of the tree's insertion-history order.  Checker of the
//              product's PRTC_ORDER_BY_ID (tree built on the device).
//
// Arithmetic goes through the same glm stand-in header as L0
// (oracle/shim/glm/glm.hpp; "parity unpinned" at the glm boundary, see there).
// Each function cites the reference file:line it restates.
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#include <atomic>
#include <thread>

#include "shim/glm/glm.hpp"

namespace {

using glm::vec3;
using glm::vec4;
using glm::mat4;
using glm::quat;

// ---------------------------------------------------------------------------
// AABB  (reference aabb.h:6-8, aabb.cpp)
// ---------------------------------------------------------------------------
struct Box {
    vec3 lo, hi;
};
// aabb.cpp:3-10  inclusive overlap, true when any compare involves NaN
inline bool box_intersect(Box const & a, Box const & b) {
    return !(a.lo.x > b.hi.x || a.hi.x < b.lo.x ||
             a.lo.y > b.hi.y || a.hi.y < b.lo.y ||
             a.lo.z > b.hi.z || a.hi.z < b.lo.z);
}
// aabb.cpp:51-58
inline bool box_contains(Box const & a, Box const & o) {
    return a.lo.x <= o.lo.x && a.lo.y <= o.lo.y && a.lo.z <= o.lo.z &&
           a.hi.x >= o.hi.x && a.hi.y >= o.hi.y && a.hi.z >= o.hi.z;
}
// aabb.cpp:60-64, aabb.h:60-63  (lhs op rhs, glm::min/max argument order matters for NaN)
inline Box box_union(Box lhs, Box const & rhs) {
    lhs.lo = glm::min(lhs.lo, rhs.lo);
    lhs.hi = glm::max(lhs.hi, rhs.hi);
    return lhs;
}
// aabb.cpp:66-69
inline float box_area(Box const & b) {
    vec3 d = b.hi - b.lo;
    return 2.0f * (d.x * d.y + d.y * d.z + d.z * d.x);
}
// aabb.cpp:13-48 (Ericson slab test)
inline bool box_intersect_ray(Box const & a, vec3 const & origin, vec3 const & direction, float maxDistance) {
    float tmin = 0.0f;
    float tmax = maxDistance;
    for (int i = 0; i < 3; i++) {
        if (std::abs(direction[i]) < 0.00001) {      // float promoted to double against a double literal
            if (origin[i] < a.lo[i] || origin[i] > a.hi[i]) return false;
        } else {
            float ood = 1.0f / direction[i];
            float t1 = (a.lo[i] - origin[i]) * ood;
            float t2 = (a.hi[i] - origin[i]) * ood;
            if (t1 > t2) { float tmp = t1; t1 = t2; t2 = tmp; }
            if (t1 > tmin) tmin = t1;
            if (t2 < tmax) tmax = t2;
            if (tmin > tmax) return false;
        }
    }
    return true;
}

enum Shape : uint8_t { SHAPE_NONE = 0, SHAPE_MODEL = 1, SHAPE_MESH = 2, SHAPE_CAPSULE = 3 };   // collider_tag.h:6-12

// ---------------------------------------------------------------------------
// prt::priority_queue  (reference src/container/priority_queue.h:19-68) -- a
// NON-standard binary min-heap whose sift-up does not stop at the first
// non-swap and whose sift-down always runs to a leaf; its tie/violation
// behaviour decides findBestSibling, so it is restated literally.
// ---------------------------------------------------------------------------
struct NodeCost {                                           // aabb_tree.h:141-147
    int32_t index;
    float directCost;
    float inheritedCost;
};
inline bool cost_less(NodeCost const & l, NodeCost const & r) {
    return l.directCost + l.inheritedCost < r.directCost + r.inheritedCost;
}
struct PrtHeap {
    std::vector<NodeCost> c;
    bool empty() const { return c.empty(); }
    NodeCost const & top() const { return c.front(); }
    void push(NodeCost const & t) {                         // priority_queue.h:19-32
        size_t index = c.size();
        c.push_back(t);
        while (index != 0) {
            size_t parentIndex = (index / 2 + (index % 2 != 0)) - 1;
            NodeCost elem = c[index];
            if (cost_less(t, c[parentIndex])) {
                c[index] = c[parentIndex];
                c[parentIndex] = elem;
            }
            index = parentIndex;
        }
    }
    void pop() {                                            // priority_queue.h:34-68
        c.front() = c.back();
        c.pop_back();
        size_t index = 0;
        while (true) {
            size_t leftIndex = index * 2 + 1;
            size_t rightIndex = leftIndex + 1;
            size_t childIndex;
            bool isLeft = leftIndex < c.size();
            bool isRight = rightIndex < c.size();
            if (isLeft && isRight) childIndex = cost_less(c[leftIndex], c[rightIndex]) ? leftIndex : rightIndex;
            else if (isLeft) childIndex = leftIndex;
            else if (isRight) childIndex = rightIndex;
            else return;
            NodeCost child = c[childIndex];
            if (cost_less(child, c[index])) {
                c[childIndex] = c[index];
                c[index] = child;
            }
            index = childIndex;
        }
    }
};

// ---------------------------------------------------------------------------
// DynamicAABBTree  (reference aabb_tree.h / aabb_tree.cpp), uint32 tags
// ---------------------------------------------------------------------------
struct Tree {
    static constexpr int32_t NIL = -1;
    struct Node {                                           // aabb_tree.h:112-139
        Box box;
        int32_t height = 0;
        int32_t parent = NIL;                               // aliases `next` in the free list
        int32_t left = NIL;
        int32_t right = NIL;
        uint32_t tagIndex = 0;
        uint8_t tagShape = SHAPE_NONE;
        bool isLeaf() const { return right == NIL; }
    };
    float buffer = 0.05f;                                   // aabb_tree.h:90
    int32_t root = NIL;
    int32_t freeHead = NIL;
    int32_t size = 0;
    std::vector<Node> nodes;
    mutable uint64_t nodeTests = 0;

    int32_t allocateNode() {                                // aabb_tree.cpp:229-241
        int32_t index;
        if (freeHead == NIL) {
            index = int32_t(nodes.size());
            nodes.push_back(Node{});
        } else {
            index = freeHead;
            freeHead = nodes[freeHead].parent;
            nodes[index] = Node{};
        }
        ++size;
        return index;
    }
    void freeNode(int32_t index) {                          // aabb_tree.cpp:222-227
        nodes[index].parent = freeHead;
        nodes[index].height = -1;
        freeHead = index;
        --size;
    }
    NodeCost siblingCost(Node const & leaf, int32_t siblingIndex, float inheritedCost) const {   // :330-339
        NodeCost cost;
        cost.index = siblingIndex;
        cost.directCost = box_area(box_union(leaf.box, nodes[siblingIndex].box));
        cost.inheritedCost = inheritedCost;
        return cost;
    }
    int32_t findBestSibling(int32_t leafIndex) const {      // aabb_tree.cpp:298-328
        Node const & leaf = nodes[leafIndex];
        int32_t bestSibling = root;
        float bestCost = box_area(box_union(leaf.box, nodes[root].box));
        PrtHeap q;
        q.push(siblingCost(leaf, root, 0.0f));
        while (!q.empty()) {
            NodeCost nc = q.top();
            q.pop();
            float currentCost = nc.directCost + nc.inheritedCost;
            if (currentCost < bestCost) {
                bestCost = currentCost;
                bestSibling = nc.index;
            }
            Node const & curr = nodes[nc.index];
            if (!curr.isLeaf()) {
                float inheritedCost = nc.inheritedCost + box_area(box_union(leaf.box, curr.box)) - box_area(curr.box);
                float cLow = box_area(leaf.box) + inheritedCost;
                if (cLow < bestCost) {
                    NodeCost left = siblingCost(leaf, curr.left, inheritedCost);
                    NodeCost right = siblingCost(leaf, curr.right, inheritedCost);
                    q.push(left);
                    q.push(right);
                }
            }
        }
        return bestSibling;
    }
    void balance(int32_t index) {                           // aabb_tree.cpp:258-296
        if (nodes[index].height < 2) return;
        int32_t ileft = nodes[index].left;
        int32_t iright = nodes[index].right;
        int32_t difference = nodes[ileft].height - nodes[iright].height;
        if (difference > 1) {
            Node & left = nodes[ileft];
            int32_t * ia = nodes[left.left].height > nodes[left.right].height ? &left.left : &left.right;
            nodes[iright].parent = ileft;
            nodes[*ia].parent = index;
            nodes[index].right = *ia;
            *ia = iright;
            left.box = box_union(nodes[left.left].box, nodes[left.right].box);
            left.height = std::max(nodes[left.left].height, nodes[left.right].height) + 1;
        } else if (difference < -1) {
            Node & right = nodes[iright];
            int32_t * ia = nodes[right.left].height > nodes[right.right].height ? &right.left : &right.right;
            nodes[ileft].parent = iright;
            nodes[*ia].parent = index;
            nodes[index].left = *ia;
            *ia = ileft;
            right.box = box_union(nodes[right.left].box, nodes[right.right].box);
            right.height = std::max(nodes[right.left].height, nodes[right.right].height) + 1;
        }
        nodes[index].height = std::max(nodes[nodes[index].left].height, nodes[nodes[index].right].height) + 1;
    }
    void synchHierarchy(int32_t index) {                    // aabb_tree.cpp:243-256
        while (index != NIL) {
            balance(index);
            int32_t left = nodes[index].left;
            int32_t right = nodes[index].right;
            nodes[index].height = 1 + std::max(nodes[left].height, nodes[right].height);
            nodes[index].box = box_union(nodes[left].box, nodes[right].box);
            index = nodes[index].parent;
        }
    }
    int32_t insertLeaf(uint32_t tagIndex, uint8_t tagShape, Box const & aabb) {   // aabb_tree.cpp:134-180
        int32_t leafIndex = allocateNode();
        nodes[leafIndex].tagIndex = tagIndex;
        nodes[leafIndex].tagShape = tagShape;
        nodes[leafIndex].box.lo = aabb.lo - buffer;
        nodes[leafIndex].box.hi = aabb.hi + buffer;
        if (root == NIL) {
            root = leafIndex;
            return leafIndex;
        }
        int32_t siblingIndex = findBestSibling(leafIndex);
        int32_t oldParentIndex = nodes[siblingIndex].parent;
        int32_t newParentIndex = allocateNode();
        nodes[newParentIndex].parent = oldParentIndex;
        nodes[newParentIndex].box = box_union(nodes[leafIndex].box, nodes[siblingIndex].box);
        nodes[newParentIndex].height = 1 + nodes[siblingIndex].height;
        if (oldParentIndex != NIL) {
            if (nodes[oldParentIndex].left == siblingIndex) nodes[oldParentIndex].left = newParentIndex;
            else nodes[oldParentIndex].right = newParentIndex;
        } else {
            root = newParentIndex;
        }
        nodes[newParentIndex].left = siblingIndex;
        nodes[newParentIndex].right = leafIndex;
        nodes[siblingIndex].parent = newParentIndex;
        nodes[leafIndex].parent = newParentIndex;
        synchHierarchy(nodes[leafIndex].parent);
        return leafIndex;
    }
    void remove(int32_t index) {                            // aabb_tree.cpp:182-220
        if (index == root) {
            root = NIL;
            freeNode(index);
            return;
        }
        int32_t parent = nodes[index].parent;
        int32_t siblingIndex = (nodes[parent].left == index) ? nodes[parent].right : nodes[parent].left;
        if (nodes[parent].parent != NIL) {
            int32_t grandParent = nodes[parent].parent;
            if (nodes[grandParent].left == parent) nodes[grandParent].left = siblingIndex;
            else nodes[grandParent].right = siblingIndex;
            nodes[siblingIndex].parent = grandParent;
            freeNode(parent);
            synchHierarchy(grandParent);
        } else {
            root = siblingIndex;
            nodes[siblingIndex].parent = NIL;
            freeNode(parent);
        }
        freeNode(index);
    }
    void update(int32_t * treeIndices, Box const * aabbs, size_t n) {   // aabb_tree.cpp:102-111
        for (size_t i = 0; i < n; ++i) {
            Node & node = nodes[treeIndices[i]];
            if (!box_contains(node.box, aabbs[i])) {
                uint32_t ti = node.tagIndex;
                uint8_t ts = node.tagShape;
                remove(treeIndices[i]);
                treeIndices[i] = insertLeaf(ti, ts, aabbs[i]);
            }
        }
    }
    // aabb_tree.cpp:34-68; caller = capsule index or UINT32_MAX; right-first DFS via explicit stack
    void query(uint32_t callerCapsule, Box const & aabb, std::vector<uint32_t> & meshIdx,
               std::vector<uint32_t> & capsuleIdx, std::vector<int32_t> & stack, uint32_t * nodeTestsOut) const {
        uint32_t tests = 0;
        if (size != 0) {
            stack.clear();
            stack.push_back(root);
            while (!stack.empty()) {
                int32_t index = stack.back();
                stack.pop_back();
                Node const & node = nodes[index];
                ++tests;
                if (box_intersect(node.box, aabb)) {
                    if (node.isLeaf()) {
                        if (node.tagShape == SHAPE_MESH) meshIdx.push_back(node.tagIndex);
                        else if (node.tagShape == SHAPE_CAPSULE && node.tagIndex != callerCapsule) capsuleIdx.push_back(node.tagIndex);
                    } else {
                        stack.push_back(node.left);
                        stack.push_back(node.right);
                    }
                }
            }
        }
        if (nodeTestsOut) *nodeTestsOut = tests;
    }
    // aabb_tree.cpp:70-93.  (A non-COLLIDE leaf would push children -1 there; every leaf here is COLLIDE.)
    void queryRaycast(vec3 const & origin, vec3 const & direction, float maxDistance, std::vector<uint32_t> & meshIdx) const {
        if (size == 0) return;
        std::vector<int32_t> stack;
        stack.push_back(root);
        while (!stack.empty()) {
            int32_t index = stack.back();
            stack.pop_back();
            Node const & node = nodes[index];
            if (box_intersect_ray(node.box, origin, direction, maxDistance)) {
                if (node.isLeaf()) {
                    if (node.tagShape == SHAPE_MESH) meshIdx.push_back(node.tagIndex);
                } else {
                    stack.push_back(node.left);
                    stack.push_back(node.right);
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------
// physics_util  (reference src/util/physics_util.cpp)
// ---------------------------------------------------------------------------
// physics_util.cpp:5-49 (Ericson, Voronoi regions)
inline vec3 closestPointOnTriangle(vec3 const & a, vec3 const & b, vec3 const & c, vec3 const & p) {
    vec3 ab = b - a;
    vec3 ac = c - a;
    vec3 bc = c - b;
    float snom = glm::dot(p - a, ab), sdenom = glm::dot(p - b, a - b);
    float tnom = glm::dot(p - a, ac), tdenom = glm::dot(p - c, a - c);
    if (snom <= 0.0f && tnom <= 0.0f) return a;
    float unom = glm::dot(p - b, bc), udenom = glm::dot(p - c, b - c);
    if (sdenom <= 0.0f && unom <= 0.0f) return b;
    if (tdenom <= 0.0f && udenom <= 0.0f) return c;
    vec3 n = glm::cross(b - a, c - a);
    float vc = glm::dot(n, glm::cross(a - p, b - p));
    if (vc <= 0.0f && snom >= 0.0f && sdenom >= 0.0f) return a + snom / (snom + sdenom) * ab;
    float va = glm::dot(n, glm::cross(b - p, c - p));
    if (va <= 0.0f && unom >= 0.0f && udenom >= 0.0f) return b + unom / (unom + udenom) * bc;
    float vb = glm::dot(n, glm::cross(c - p, a - p));
    if (vb <= 0.0f && tnom >= 0.0f && tdenom >= 0.0f) return a + tnom / (tnom + tdenom) * ac;
    float u = va / (va + vb + vc);
    float v = vb / (va + vb + vc);
    float w = 1.0f - u - v;
    return u * a + v * b + w * c;
}
// physics_util.cpp:125-139 (Nettle)
inline vec3 closestPointOnLineSegment(vec3 const & a, vec3 const & b, vec3 const & p) {
    vec3 c = p - a;
    vec3 v = glm::normalize(b - a);
    float d = glm::distance(a, b);
    float t = glm::dot(v, c);
    if (t < 0) return a;
    if (t > d) return b;
    v = v * t;
    return a + v;
}
// physics_util.cpp:181-216 (Ericson)
inline bool intersectLineSegmentTriangle(vec3 const & origin, vec3 const & end, vec3 const & a, vec3 const & b,
                                         vec3 const & c, float & t) {
    vec3 ab = b - a;
    vec3 ac = c - a;
    vec3 qp = origin - end;
    vec3 n = glm::cross(ab, ac);
    float d = glm::dot(qp, n);
    if (d <= 0.0f) return false;
    vec3 ap = origin - a;
    t = glm::dot(ap, n);
    if (t < 0.0f) return false;
    if (t > d) return false;
    vec3 e = glm::cross(qp, ap);
    float v = glm::dot(ac, e);
    if (v < 0.0f || v > d) return false;
    float w = -glm::dot(ab, e);
    if (w < 0.0f || v + w > d) return false;
    float ood = 1.0f / d;
    t *= ood;
    return true;
}

// ---------------------------------------------------------------------------
// world
// ---------------------------------------------------------------------------
struct Config {                  // mirrors prtc_config (include/prtc.h)
    float gravity = 1.0f;        // physics_system.h:122
    uint32_t substeps = 4;       // physics_system.cpp:346
    float leaf_margin = 0.05f;   // aabb_tree.h:90
    float ground_dot = 0.3f;     // collision_system.cpp:245
    float friction = 10.0f;      // physics_system.cpp:307
    float stick = 0.05f;         // physics_system.cpp:280
    uint32_t abs_mode = 0;       // 0: `abs(dist)` -> int abs(int) (libstdc++ build), 1: float fabs (libc++ build)
    uint32_t order_mode = 0;     // 0 LITERAL (unified tree, sequential), 1 CANONICAL (static-only tree), 2 BY_ID (as 1, mesh candidates sorted by id)
    uint32_t dyn_mode = 0;       // CANONICAL only: 0 OFF, 2 JACOBI
    uint32_t threads = 1;        // CANONICAL only (std::thread over capsules)
};

struct Capsule { float height, radius; vec3 offset; };                  // colliders.h:14-18
struct Transform { vec3 position; quat rotation; vec3 scale; };         // component.h:14-17

inline mat4 capsuleTform(vec3 const & position, quat const & rotation) {   // physics_system.cpp:267,344,352; collision_system.cpp:40
    return glm::translate(mat4(1.0f), position) * glm::toMat4(glm::normalize(rotation));
}
inline mat4 transformMatrix(Transform const & t) {                      // component.h:19-24
    mat4 scaleM = glm::scale(t.scale);
    mat4 rotateM = glm::toMat4(glm::normalize(t.rotation));
    mat4 translateM = glm::translate(mat4(1.0f), t.position);
    return translateM * rotateM * scaleM;
}
inline Box capsuleAABB(Capsule const & c, mat4 const & transform) {     // colliders.h:19-27
    vec4 a{c.offset, 1.0f};
    vec4 b{c.offset + vec3{0.0f, c.height, 0.0f}, 1.0f};
    vec3 tA = transform * a;
    vec3 tB = transform * b;
    return { glm::min(tA, tB) - vec3{c.radius}, glm::max(tA, tB) + vec3{c.radius} };
}

// `abs(dist)` at collision_system.cpp:72 is UNQUALIFIED.  Under libstdc++ only `int abs(int)` is visible at
// global scope there, so dist is truncated to int (x86 cvttss2si: out-of-range/NaN -> INT_MIN, and
// abs(INT_MIN) stays INT_MIN); under libc++ the float overload is chosen.  Both are restated.
inline float absDist(float dist, uint32_t abs_mode) {
    if (abs_mode == 1) return std::fabs(dist);
    int di;
    if (!(dist < 2147483648.0f) || !(dist >= -2147483648.0f)) di = INT_MIN;
    else di = int(dist);
    unsigned ua = di < 0 ? 0u - unsigned(di) : unsigned(di);
    return float(int(ua));
}

struct Trace {                   // same record layout as oracle/l0_driver.cpp
    bool on = false;
    std::vector<uint32_t> q_char, q_sub;
    std::vector<float> q_aabb;
    std::vector<uint32_t> q_mesh_off, q_caps_off, mesh_ids, caps_ids, q_tri_tests, q_node_tests;
    std::vector<uint32_t> h_query;
    std::vector<int32_t> h_poly;
    std::vector<float> h_impulse, h_normal, h_pos;
    void clear() {
        q_char.clear(); q_sub.clear(); q_aabb.clear();
        q_mesh_off.assign(1, 0); q_caps_off.assign(1, 0);
        mesh_ids.clear(); caps_ids.clear(); q_tri_tests.clear(); q_node_tests.clear();
        h_query.clear(); h_poly.clear(); h_impulse.clear(); h_normal.clear(); h_pos.clear();
    }
};

struct Counters { uint64_t substeps = 0, tri_tests = 0, node_tests = 0, contacts = 0, cc_tests = 0; };

struct MeshCol { uint32_t first; uint32_t count; uint32_t model; };    // colliders.h:38-47 (first/count in vertices of `cache`)

struct World {
    Config cfg;
    // static geometry (physics_system.h:83-101)
    std::vector<vec3> raw, cache;            // all models back to back
    std::vector<MeshCol> meshes;
    std::vector<Box> meshAABBs;
    std::vector<int32_t> meshTreeIdx;
    struct ModelCol { uint32_t firstMesh, numMeshes; };
    std::vector<ModelCol> models;
    // capsules / characters (character i <-> capsule i)
    std::vector<Capsule> capsules;
    std::vector<Box> capsuleAABBs;
    std::vector<int32_t> capsuleTreeIdx;     // LITERAL
    std::vector<Box> capsuleStored;          // CANONICAL: stored fat AABB (hysteresis, aabb_tree.cpp:102-111,140-141)
    std::vector<Transform> transforms;
    std::vector<vec3> velocity, movement, groundNormal;
    std::vector<uint8_t> grounded;
    Tree tree;
    Trace trace;
    Counters counters;
    // JACOBI neighbour search at scale: uniform xz grid over the stored fat boxes, rebuilt once per frame (the stored boxes
    // are fixed within a frame).  Pure acceleration: a query collects the boxes binned in the cells it touches, de-duplicates,
    // keeps those that pass the same box_intersect as the brute-force loop, ascending.  Boxes or queries with a non-finite
    // coordinate (box_intersect is true for NaN, aabb.cpp:3-10) bypass the grid.
    struct DynGrid {
        float ox = 0, oz = 0, inv = 0; int gx = 0, gz = 0;
        std::vector<uint32_t> start, items, always;
        bool valid = false;
    } dyn_grid;
};

struct Scratch {
    std::vector<uint32_t> meshIDs, capsIDs;
    std::vector<int32_t> stack;
};

// collision_system.cpp:236-265 (COLLIDE branch; the event hash-sets have no reader and garbage keys -> not restated)
inline void handleCollision(World & w, uint32_t ci, vec3 const & impulse, vec3 const & normal, int32_t poly, Counters & cnt) {
    w.transforms[ci].position += impulse;
    ++cnt.contacts;
    if (w.trace.on) {
        Trace & t = w.trace;
        t.h_query.push_back(uint32_t(t.q_char.size() - 1));
        t.h_poly.push_back(poly);
        for (int k = 0; k < 3; ++k) t.h_impulse.push_back(impulse[k]);
        for (int k = 0; k < 3; ++k) t.h_normal.push_back(normal[k]);
        for (int k = 0; k < 3; ++k) t.h_pos.push_back(w.transforms[ci].position[k]);
    }
    bool groundCollision = glm::dot(normal, vec3{0.0f, 1.0f, 0.0f}) > w.cfg.ground_dot;
    if (groundCollision) {
        w.grounded[ci] = 1;
        w.groundNormal[ci] = normal;
    }
}

// collision_system.cpp:29-155, one triangle (a,b,c) against character ci at its CURRENT position.
inline void collideCapsuleTriangle(World & w, uint32_t ci, Capsule const & capsule, vec3 const & pa, vec3 const & pb,
                                   vec3 const & pc, int32_t poly, Counters & cnt) {
    Transform & transform = w.transforms[ci];
    vec4 a{capsule.offset, 1.0f};
    vec4 b{capsule.offset + vec3{0.0f, capsule.height, 0.0f}, 1.0f};
    mat4 tform = capsuleTform(transform.position, transform.rotation);
    vec3 A = tform * a;
    vec3 B = tform * b;
    vec3 CapsuleNormal = glm::normalize(A - B);
    vec3 LineEndOffset = CapsuleNormal * capsule.radius;
    vec3 pointA = A - LineEndOffset;

    vec3 n = glm::cross(pb - pa, pc - pa);
    if (glm::length(n) == 0.0f) return;
    n = glm::normalize(n);

    float nDotCn = glm::dot(n, CapsuleNormal);
    vec3 reference_point;
    if (glm::abs(nDotCn) > 0.0001f) {
        float t = glm::dot(n, (pa - pointA) / glm::abs(nDotCn));
        vec3 line_plane_intersection = pointA + CapsuleNormal * t;
        reference_point = closestPointOnTriangle(pa, pb, pc, line_plane_intersection);
    } else {
        reference_point = pa;
    }
    vec3 center = closestPointOnLineSegment(A, B, reference_point);

    float dist = glm::dot(center - pa, n);
    if (dist < 0.0f) return;
    if (absDist(dist, w.cfg.abs_mode) > capsule.radius) return;

    vec3 point0 = center - n * dist;
    vec3 c0 = glm::cross(point0 - pa, pb - pa);
    vec3 c1 = glm::cross(point0 - pb, pc - pb);
    vec3 c2 = glm::cross(point0 - pc, pa - pc);
    bool inside = glm::dot(c0, n) <= 0 && glm::dot(c1, n) <= 0 && glm::dot(c2, n) <= 0;

    float radiussq = capsule.radius * capsule.radius;
    vec3 point1 = closestPointOnLineSegment(pa, pb, center);
    vec3 v1 = center - point1;
    float distsq1 = glm::dot(v1, v1);
    bool intersects = distsq1 < radiussq;
    vec3 point2 = closestPointOnLineSegment(pb, pc, center);
    vec3 v2 = center - point2;
    float distsq2 = glm::dot(v2, v2);
    intersects |= distsq2 < radiussq;
    vec3 point3 = closestPointOnLineSegment(pc, pa, center);
    vec3 v3 = center - point3;
    float distsq3 = glm::dot(v3, v3);
    intersects |= distsq3 < radiussq;

    if (inside || intersects) {
        vec3 intersection_vec;
        if (inside) {
            intersection_vec = center - point0;
        } else {
            vec3 d = center - point1;
            float best_distsq = glm::dot(d, d);
            intersection_vec = d;
            d = center - point2;
            float distsq = glm::dot(d, d);
            if (distsq < best_distsq) {          // collision_system.cpp:115-118: `distsq = best_distsq;` (sic) --
                intersection_vec = d;            // best_distsq never shrinks
            }
            d = center - point3;
            distsq = glm::dot(d, d);
            if (distsq < best_distsq) {          // :123-127, same typo
                intersection_vec = d;
            }
        }
        float len = glm::length(intersection_vec);
        vec3 impulse{0.0f};
        vec3 normal;
        if (len > 0.0f && len < capsule.radius) {
            float penetration_depth = capsule.radius - len;
            vec3 penetration_normal = intersection_vec / len;
            impulse = penetration_depth * (penetration_normal + 0.0001f);
            normal = penetration_normal;
        } else {
            normal = n;
        }
        handleCollision(w, ci, impulse, normal, poly, cnt);
    }
}

// collision_system.cpp:158-234; pose of B is passed explicitly (LITERAL: its live transform; JACOBI: snapshot)
inline void collideCapsuleCapsule(World & w, uint32_t ia, Capsule const & capsuleA, vec3 const & posB, quat const & rotB,
                                  Capsule const & capsuleB, Counters & cnt) {
    vec4 aa{capsuleA.offset, 1.0f};
    vec4 ab{capsuleA.offset + vec3{0.0f, capsuleA.height, 0.0f}, 1.0f};
    Transform & transformA = w.transforms[ia];
    mat4 tformA = capsuleTform(transformA.position, transformA.rotation);
    vec3 a_A = tformA * aa;
    vec3 a_B = tformA * ab;
    vec4 ba{capsuleB.offset, 1.0f};
    vec4 bb{capsuleB.offset + vec3{0.0f, capsuleB.height, 0.0f}, 1.0f};
    mat4 tformB = capsuleTform(posB, rotB);
    vec3 b_A = tformB * ba;
    vec3 b_B = tformB * bb;
    vec3 v0 = b_A - a_A;
    vec3 v1 = b_B - a_A;
    vec3 v2 = b_A - a_B;
    vec3 v3 = b_B - a_B;
    float d0 = glm::dot(v0, v0);
    float d1 = glm::dot(v1, v1);
    float d2 = glm::dot(v2, v2);
    float d3 = glm::dot(v3, v3);
    vec3 bestA;
    if (d2 < d0 || d2 < d1 || d3 < d0 || d3 < d1) bestA = a_B;
    else bestA = a_A;
    vec3 bestB = closestPointOnLineSegment(b_A, b_B, bestA);
    bestA = closestPointOnLineSegment(a_A, a_B, bestB);
    vec3 penetration_normal = bestA - bestB;
    float len = glm::length(penetration_normal);
    penetration_normal /= len;                       // len == 0 -> NaN, unguarded (:210-211)
    float penetration_depth = capsuleA.radius + capsuleB.radius - len;
    bool intersects = penetration_depth > 0;
    ++cnt.cc_tests;
    if (intersects) {
        vec3 impulse = penetration_depth * (penetration_normal + 0.0001f);
        handleCollision(w, ia, impulse, penetration_normal, -1, cnt);
    }
}

// physics_system.cpp:330-438
void collideCharacterWithWorld(World & w, uint32_t ci, Scratch & s, Counters & cnt, std::vector<Transform> const * snapshot) {
    Capsule const & capsule = w.capsules[ci];
    Transform & transform = w.transforms[ci];
    vec3 const velocity = w.velocity[ci];
    mat4 prevTform = capsuleTform(transform.position, transform.rotation);
    const uint32_t nTimeSteps = w.cfg.substeps;
    uint32_t stepsLeft = nTimeSteps;
    while (stepsLeft != 0) {
        --stepsLeft;
        transform.position += velocity / float(nTimeSteps);
        mat4 tform = capsuleTform(transform.position, transform.rotation);
        Box eAABB = capsuleAABB(capsule, prevTform);
        eAABB = box_union(eAABB, capsuleAABB(capsule, tform));
        w.capsuleAABBs[ci] = eAABB;
        prevTform = tform;

        s.meshIDs.clear();
        s.capsIDs.clear();
        uint32_t nodeTests = 0;
        if (w.cfg.order_mode == 0) {
            w.tree.query(ci, eAABB, s.meshIDs, s.capsIDs, s.stack, &nodeTests);
        } else {
            w.tree.query(UINT32_MAX, eAABB, s.meshIDs, s.capsIDs, s.stack, &nodeTests);
            // BY_ID: the candidate SET of the reference's query (it does not depend on the tree's shape), taken in
            // ascending collider id = registration order instead of the insertion-history order of the tree
            if (w.cfg.order_mode == 2) std::sort(s.meshIDs.begin(), s.meshIDs.end());
            if (w.cfg.dyn_mode == 2) {
                // canonical neighbour rule: ascending capsule index over the stored fat AABBs
                World::DynGrid const & g = w.dyn_grid;
                const float fin = (eAABB.lo.x + eAABB.lo.z) + (eAABB.hi.x + eAABB.hi.z);
                if (!g.valid || !(fin - fin == 0.0f)) {
                    for (uint32_t j = 0; j < uint32_t(w.capsules.size()); ++j)
                        if (j != ci && box_intersect(w.capsuleStored[j], eAABB)) s.capsIDs.push_back(j);
                } else {
                    auto cell = [&](float v, float o, int n) { int c = int(std::floor((v - o) * g.inv)); return c < 0 ? 0 : (c >= n ? n - 1 : c); };
                    const int x0 = cell(eAABB.lo.x, g.ox, g.gx), x1 = cell(eAABB.hi.x, g.ox, g.gx);
                    const int z0 = cell(eAABB.lo.z, g.oz, g.gz), z1 = cell(eAABB.hi.z, g.oz, g.gz);
                    for (int z = z0; z <= z1; ++z)
                        for (int x = x0; x <= x1; ++x)
                            for (uint32_t k = g.start[size_t(z) * g.gx + x]; k < g.start[size_t(z) * g.gx + x + 1]; ++k) s.capsIDs.push_back(g.items[k]);
                    for (uint32_t j : g.always) s.capsIDs.push_back(j);
                    std::sort(s.capsIDs.begin(), s.capsIDs.end());
                    s.capsIDs.erase(std::unique(s.capsIDs.begin(), s.capsIDs.end()), s.capsIDs.end());
                    size_t keep = 0;
                    for (uint32_t j : s.capsIDs)
                        if (j != ci && box_intersect(w.capsuleStored[j], eAABB)) s.capsIDs[keep++] = j;
                    s.capsIDs.resize(keep);
                }
            }
        }
        ++cnt.substeps;
        cnt.node_tests += nodeTests;
        if (w.trace.on) {
            Trace & t = w.trace;
            t.q_char.push_back(ci);
            t.q_sub.push_back(nTimeSteps - 1 - stepsLeft);
            t.q_aabb.push_back(eAABB.lo.x); t.q_aabb.push_back(eAABB.lo.y); t.q_aabb.push_back(eAABB.lo.z);
            t.q_aabb.push_back(eAABB.hi.x); t.q_aabb.push_back(eAABB.hi.y); t.q_aabb.push_back(eAABB.hi.z);
            for (uint32_t id : s.meshIDs) t.mesh_ids.push_back(id);
            for (uint32_t id : s.capsIDs) t.caps_ids.push_back(id);
            t.q_mesh_off.push_back(uint32_t(t.mesh_ids.size()));
            t.q_caps_off.push_back(uint32_t(t.caps_ids.size()));
            t.q_tri_tests.push_back(0);
            t.q_node_tests.push_back(nodeTests);
        }

        for (uint32_t colID : s.capsIDs) {                               // physics_system.cpp:369-389
            Transform const & other = snapshot ? (*snapshot)[colID] : w.transforms[colID];
            collideCapsuleCapsule(w, ci, capsule, other.position, other.rotation, w.capsules[colID], cnt);
        }
        if (s.meshIDs.empty()) break;                                    // :391-393

        uint32_t poly = 0;                                               // :396-434 without the aggregate copy
        for (uint32_t colID : s.meshIDs) {
            MeshCol const & m = w.meshes[colID];
            for (uint32_t v = m.first; v + 2 < m.first + m.count; v += 3) {
                ++cnt.tri_tests;
                if (w.trace.on) ++w.trace.q_tri_tests.back();
                collideCapsuleTriangle(w, ci, capsule, w.cache[v], w.cache[v + 1], w.cache[v + 2], int32_t(poly), cnt);
                ++poly;
            }
        }
    }
    transform.position += float(stepsLeft) * velocity / float(nTimeSteps);   // :437
}

// physics_system.cpp:259-288 for one character
inline void phase1(World & w, uint32_t i, float deltaTime, vec3 & prevVelocity) {
    Capsule const & capsule = w.capsules[i];
    Transform & transform = w.transforms[i];
    mat4 tform = capsuleTform(transform.position, transform.rotation);
    Box eAABB = capsuleAABB(capsule, tform);
    float gravityFactor = w.cfg.gravity;
    mat4 velTform = glm::translate(tform, w.velocity[i] + vec3{0.0f, -1.0f, 0.0f} * gravityFactor * deltaTime);
    eAABB = box_union(eAABB, capsuleAABB(capsule, velTform));
    w.capsuleAABBs[i] = eAABB;
    prevVelocity = w.velocity[i];
    if (w.grounded[i]) {
        w.velocity[i] += (-w.cfg.stick * gravityFactor * deltaTime) * w.groundNormal[i];
    }
    w.velocity[i].x += w.movement[i].x;
    w.velocity[i].z += w.movement[i].z;
    w.grounded[i] = 0;
}
// physics_system.cpp:298-317 for one character
inline void phase3(World & w, uint32_t i, float deltaTime, vec3 const & prevVelocity) {
    float gravityFactor = w.cfg.gravity;
    w.velocity[i] = prevVelocity;
    float frictionRatio = 1 / (1 + (deltaTime * w.cfg.friction));
    w.velocity[i].x = w.velocity[i].x * frictionRatio;
    w.velocity[i].z = w.velocity[i].z * frictionRatio;
    if (w.grounded[i]) {
        w.velocity[i].y = glm::max(0.0f * gravityFactor * deltaTime, w.velocity[i].y);
    } else {
        w.velocity[i].y += -1.0f * gravityFactor * deltaTime;
    }
}

// physics_system.cpp:245-318
void updateCharacters(World & w, float deltaTime) {
    const uint32_t n = uint32_t(w.capsules.size());
    std::vector<vec3> prevVelocities(n);
    for (uint32_t i = 0; i < n; ++i) phase1(w, i, deltaTime, prevVelocities[i]);

    if (w.cfg.order_mode == 0) {
        // LITERAL: refresh every capsule leaf in the unified tree, then collide sequentially
        w.tree.update(w.capsuleTreeIdx.data(), w.capsuleAABBs.data(), n);
        Scratch s;
        for (uint32_t i = 0; i < n; ++i) collideCharacterWithWorld(w, i, s, w.counters, nullptr);
    } else {
        if (w.cfg.dyn_mode == 2) {
            for (uint32_t i = 0; i < n; ++i) {          // stored fat AABB hysteresis (aabb_tree.cpp:102-111,140-141)
                if (!box_contains(w.capsuleStored[i], w.capsuleAABBs[i])) {
                    w.capsuleStored[i].lo = w.capsuleAABBs[i].lo - w.cfg.leaf_margin;
                    w.capsuleStored[i].hi = w.capsuleAABBs[i].hi + w.cfg.leaf_margin;
                }
            }
        }
        if (w.cfg.dyn_mode == 0) {
            // independent units: capsules are partitioned over host threads in chunks of 256
            int threads = w.trace.on ? 1 : int(std::max(1u, w.cfg.threads));
            std::vector<Counters> tc{static_cast<size_t>(threads)};
            std::atomic<uint32_t> next{0};
            auto worker = [&](int tid) {
                Scratch s;
                Counters & c = tc[size_t(tid)];
                for (;;) {
                    uint32_t begin = next.fetch_add(256u);
                    if (begin >= n) break;
                    uint32_t end = std::min(n, begin + 256u);
                    for (uint32_t i = begin; i < end; ++i) collideCharacterWithWorld(w, i, s, c, nullptr);
                }
            };
            if (threads == 1) {
                worker(0);
            } else {
                std::vector<std::thread> pool;
                for (int t = 0; t < threads; ++t) pool.emplace_back(worker, t);
                for (auto & th : pool) th.join();
            }
            for (auto & c : tc) {
                w.counters.substeps += c.substeps; w.counters.tri_tests += c.tri_tests;
                w.counters.node_tests += c.node_tests; w.counters.contacts += c.contacts; w.counters.cc_tests += c.cc_tests;
            }
        } else {
            // JACOBI (frame-level snapshot): every capsule collides against the start-of-frame pose of the others
            std::vector<Transform> snapshot(w.transforms);
            {   // the grid over this frame's stored boxes (every box in every cell it touches)
                World::DynGrid & g = w.dyn_grid;
                g.valid = false; g.always.clear();
                float lox = 3e38f, loz = 3e38f, hix = -3e38f, hiz = -3e38f;
                std::vector<uint8_t> finite(n);
                for (uint32_t i = 0; i < n; ++i) {
                    Box const & b = w.capsuleStored[i];
                    const float f = (b.lo.x + b.lo.z) + (b.hi.x + b.hi.z);
                    finite[i] = (f - f == 0.0f) && std::fabs(b.lo.x) < 1e7f && std::fabs(b.hi.x) < 1e7f && std::fabs(b.lo.z) < 1e7f && std::fabs(b.hi.z) < 1e7f;
                    if (!finite[i]) { g.always.push_back(i); continue; }
                    lox = std::min(lox, b.lo.x); loz = std::min(loz, b.lo.z); hix = std::max(hix, b.hi.x); hiz = std::max(hiz, b.hi.z);
                }
                if (n >= 2048 && hix > lox && hiz > loz && !std::getenv("PRT_L1_NO_GRID")) {     // (the switch lets the tests compare with the plain loop)
                    const float cellsz = std::max(4.0f, std::max(hix - lox, hiz - loz) / 1024.0f);
                    g.ox = lox; g.oz = loz; g.inv = 1.0f / cellsz;
                    g.gx = std::max(1, int(std::ceil((hix - lox) * g.inv)) + 1); g.gz = std::max(1, int(std::ceil((hiz - loz) * g.inv)) + 1);
                    auto cell = [&](float v, float o, int m) { int c = int(std::floor((v - o) * g.inv)); return c < 0 ? 0 : (c >= m ? m - 1 : c); };
                    g.start.assign(size_t(g.gx) * g.gz + 1, 0u);
                    for (int pass = 0; pass < 2; ++pass) {
                        std::vector<uint32_t> fill(pass ? g.start : std::vector<uint32_t>());
                        for (uint32_t i = 0; i < n; ++i) {
                            if (!finite[i]) continue;
                            Box const & b = w.capsuleStored[i];
                            for (int z = cell(b.lo.z, g.oz, g.gz); z <= cell(b.hi.z, g.oz, g.gz); ++z)
                                for (int x = cell(b.lo.x, g.ox, g.gx); x <= cell(b.hi.x, g.ox, g.gx); ++x) {
                                    if (!pass) ++g.start[size_t(z) * g.gx + x + 1];
                                    else g.items[fill[size_t(z) * g.gx + x]++] = i;
                                }
                        }
                        if (!pass) {
                            for (size_t c = 1; c < g.start.size(); ++c) g.start[c] += g.start[c - 1];
                            g.items.assign(g.start.back(), 0u);
                        }
                    }
                    g.valid = true;
                }
            }
            // every capsule reads the snapshot and its own state only: independent units, like the static pass
            int threads = w.trace.on ? 1 : int(std::max(1u, w.cfg.threads));
            std::vector<Counters> tc{static_cast<size_t>(threads)};
            std::atomic<uint32_t> next{0};
            auto worker = [&](int tid) {
                Scratch s;
                Counters & c = tc[size_t(tid)];
                for (;;) {
                    uint32_t begin = next.fetch_add(256u);
                    if (begin >= n) break;
                    uint32_t end = std::min(n, begin + 256u);
                    for (uint32_t i = begin; i < end; ++i) collideCharacterWithWorld(w, i, s, c, &snapshot);
                }
            };
            if (threads == 1) {
                worker(0);
            } else {
                std::vector<std::thread> pool;
                for (int t = 0; t < threads; ++t) pool.emplace_back(worker, t);
                for (auto & th : pool) th.join();
            }
            for (auto & c : tc) {
                w.counters.substeps += c.substeps; w.counters.tri_tests += c.tri_tests;
                w.counters.node_tests += c.node_tests; w.counters.contacts += c.contacts; w.counters.cc_tests += c.cc_tests;
            }
            w.dyn_grid.valid = false;
        }
    }
    for (uint32_t i = 0; i < n; ++i) phase3(w, i, deltaTime, prevVelocities[i]);
}

} // namespace

// ---------------------------------------------------------------------------
// C-ABI (ctypes)
// ---------------------------------------------------------------------------
extern "C" {

void * l1_create(const float * cfg_f /*gravity, leaf_margin, ground_dot, friction, stick*/,
                 const uint32_t * cfg_u /*substeps, abs_mode, order_mode, dyn_mode, threads*/) {
    World * w = new World();
    if (cfg_f) { w->cfg.gravity = cfg_f[0]; w->cfg.leaf_margin = cfg_f[1]; w->cfg.ground_dot = cfg_f[2]; w->cfg.friction = cfg_f[3]; w->cfg.stick = cfg_f[4]; }
    if (cfg_u) { w->cfg.substeps = cfg_u[0]; w->cfg.abs_mode = cfg_u[1]; w->cfg.order_mode = cfg_u[2]; w->cfg.dyn_mode = cfg_u[3]; w->cfg.threads = cfg_u[4]; }
    w->tree.buffer = w->cfg.leaf_margin;
    return w;
}
void l1_destroy(void * h) { delete static_cast<World *>(h); }

// physics_system.cpp:44-106
int l1_add_model(void * h, const float * xyz, uint32_t n_idx, const uint32_t * mesh_first, const uint32_t * mesh_count,
                 uint32_t n_meshes, const float * tform10) {
    World & w = *static_cast<World *>(h);
    Transform t;
    t.position = vec3(tform10[0], tform10[1], tform10[2]);
    t.rotation = quat(tform10[3], tform10[4], tform10[5], tform10[6]);
    t.scale = vec3(tform10[7], tform10[8], tform10[9]);
    uint32_t modelIndex = uint32_t(w.models.size());
    w.models.push_back({uint32_t(w.meshes.size()), n_meshes});
    uint32_t base = uint32_t(w.cache.size());
    w.raw.resize(base + n_idx);
    w.cache.resize(base + n_idx);
    uint32_t i = 0;
    for (uint32_t m = 0; m < n_meshes; ++m) {
        uint32_t index = mesh_first[m];
        uint32_t endIndex = index + mesh_count[m];
        w.meshes.push_back({base + i, mesh_count[m], modelIndex});
        vec3 mn = vec3(std::numeric_limits<float>::max());
        vec3 mx = vec3(std::numeric_limits<float>::lowest());
        mat4 mat = transformMatrix(t);
        while (index < endIndex) {
            w.raw[base + i] = vec3(xyz[3 * index], xyz[3 * index + 1], xyz[3 * index + 2]);
            w.cache[base + i] = mat * vec4(w.raw[base + i], 1.0f);
            mn = glm::min(mn, w.cache[base + i]);
            mx = glm::max(mx, w.cache[base + i]);
            ++i;
            ++index;
        }
        w.meshAABBs.push_back({mn, mx});
    }
    size_t prevSize = w.meshTreeIdx.size();
    w.meshTreeIdx.resize(prevSize + n_meshes);
    for (size_t k = prevSize; k < prevSize + n_meshes; ++k)
        w.meshTreeIdx[k] = w.tree.insertLeaf(uint32_t(k), SHAPE_MESH, w.meshAABBs[k]);
    return int(modelIndex);
}

// physics_system.cpp:161-202 for one (tag, transform) pair: world cache and mesh boxes of the model recomputed, then
// DynamicAABBTree::update over the model's leaves (a leaf keeps its stored fat box while it still contains the new one)
void l1_update_model(void * h, uint32_t model_index, const float * tform10) {
    World & w = *static_cast<World *>(h);
    Transform t;
    t.position = vec3(tform10[0], tform10[1], tform10[2]);
    t.rotation = quat(tform10[3], tform10[4], tform10[5], tform10[6]);
    t.scale = vec3(tform10[7], tform10[8], tform10[9]);
    World::ModelCol const & col = w.models[model_index];
    mat4 mat = transformMatrix(t);
    for (uint32_t m = col.firstMesh; m < col.firstMesh + col.numMeshes; ++m) {
        MeshCol const & curr = w.meshes[m];
        vec3 mn = vec3(std::numeric_limits<float>::max());
        vec3 mx = vec3(std::numeric_limits<float>::lowest());
        for (uint32_t v = curr.first; v < curr.first + curr.count; ++v) {
            w.cache[v] = mat * vec4(w.raw[v], 1.0f);
            mn = glm::min(mn, w.cache[v]);
            mx = glm::max(mx, w.cache[v]);
        }
        w.meshAABBs[m] = {mn, mx};
    }
    w.tree.update(&w.meshTreeIdx[col.firstMesh], &w.meshAABBs[col.firstMesh], col.numMeshes);
}

// PhysicsSystem::removeCollider for a MODEL tag (physics_system.cpp:108-149): the model's mesh leaves leave the tree, in
// mesh order (aabb_tree.cpp: remove per leaf).  The widened restatement keeps the ids of the other colliders stable -- the
// reference compacts its mesh arrays but leaves the tree's tags stale (:141-146), which only works out when the removed
// model is the last one (the envelope in which L0 and this function are compared, tests/test_oracle_pinning.py); the
// follow-up tree.update of the later leaves (:137-140) is a no-op (their boxes have not changed).
void l1_remove_model(void * h, uint32_t model_index) {
    World & w = *static_cast<World *>(h);
    World::ModelCol & col = w.models[model_index];
    for (uint32_t m = col.firstMesh; m < col.firstMesh + col.numMeshes; ++m) {
        if (w.meshTreeIdx[m] >= 0) w.tree.remove(w.meshTreeIdx[m]);
        w.meshTreeIdx[m] = -1;
        w.meshes[m].count = 0;
    }
    col.numMeshes = 0;
}

// physics_system.cpp:22-42 (+ the character that owns the capsule)
int l1_add_capsule(void * h, float height, float radius, const float * off3) {
    World & w = *static_cast<World *>(h);
    uint32_t id = uint32_t(w.capsules.size());
    Capsule c{height, radius, vec3(off3[0], off3[1], off3[2])};
    w.capsules.push_back(c);
    Box ident = capsuleAABB(c, mat4{1.0f});
    w.capsuleAABBs.push_back(ident);
    if (w.cfg.order_mode == 0) {
        w.capsuleTreeIdx.push_back(w.tree.insertLeaf(id, SHAPE_CAPSULE, ident));
    } else {
        Box stored;
        stored.lo = ident.lo - w.cfg.leaf_margin;
        stored.hi = ident.hi + w.cfg.leaf_margin;
        w.capsuleStored.push_back(stored);
    }
    Transform t;
    t.position = vec3(0.0f); t.rotation = quat(1.0f, 0.0f, 0.0f, 0.0f); t.scale = vec3(1.0f);
    w.transforms.push_back(t);
    w.velocity.push_back(vec3(0.0f)); w.movement.push_back(vec3(0.0f)); w.groundNormal.push_back(vec3(0.0f));
    w.grounded.push_back(0);
    return int(id);
}

uint32_t l1_num_characters(void * h) { return uint32_t(static_cast<World *>(h)->capsules.size()); }

void l1_set_state(void * h, const float * pos3, const float * quat4_wxyz, const float * vel3, const float * move3,
                  const float * gnormal3, const uint8_t * grounded) {
    World & w = *static_cast<World *>(h);
    size_t n = w.capsules.size();
    for (size_t i = 0; i < n; ++i) {
        if (pos3) w.transforms[i].position = vec3(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]);
        if (quat4_wxyz) w.transforms[i].rotation = quat(quat4_wxyz[4 * i], quat4_wxyz[4 * i + 1], quat4_wxyz[4 * i + 2], quat4_wxyz[4 * i + 3]);
        if (vel3) w.velocity[i] = vec3(vel3[3 * i], vel3[3 * i + 1], vel3[3 * i + 2]);
        if (move3) w.movement[i] = vec3(move3[3 * i], move3[3 * i + 1], move3[3 * i + 2]);
        if (gnormal3) w.groundNormal[i] = vec3(gnormal3[3 * i], gnormal3[3 * i + 1], gnormal3[3 * i + 2]);
        if (grounded) w.grounded[i] = grounded[i] != 0;
    }
}
void l1_get_state(void * h, float * pos3, float * vel3, float * gnormal3, uint8_t * grounded) {
    World & w = *static_cast<World *>(h);
    size_t n = w.capsules.size();
    for (size_t i = 0; i < n; ++i) {
        if (pos3) { pos3[3 * i] = w.transforms[i].position.x; pos3[3 * i + 1] = w.transforms[i].position.y; pos3[3 * i + 2] = w.transforms[i].position.z; }
        if (vel3) { vel3[3 * i] = w.velocity[i].x; vel3[3 * i + 1] = w.velocity[i].y; vel3[3 * i + 2] = w.velocity[i].z; }
        if (gnormal3) { gnormal3[3 * i] = w.groundNormal[i].x; gnormal3[3 * i + 1] = w.groundNormal[i].y; gnormal3[3 * i + 2] = w.groundNormal[i].z; }
        if (grounded) grounded[i] = w.grounded[i];
    }
}

void l1_step(void * h, float dt, int trace_on) {
    World & w = *static_cast<World *>(h);
    w.trace.on = trace_on != 0;
    w.trace.clear();
    updateCharacters(w, dt);
}

void l1_counters(void * h, uint64_t * out5 /*substeps, tri_tests, node_tests, contacts, cc_tests*/, int reset) {
    World & w = *static_cast<World *>(h);
    out5[0] = w.counters.substeps; out5[1] = w.counters.tri_tests; out5[2] = w.counters.node_tests;
    out5[3] = w.counters.contacts; out5[4] = w.counters.cc_tests;
    if (reset) w.counters = Counters{};
}

// right-first preorder (= explicit-stack DFS order of aabb_tree.cpp:34-68) of all leaves
void l1_dump_order(void * h, uint32_t * mesh_ids, uint32_t * n_mesh, uint32_t * caps_ids, uint32_t * n_caps) {
    World & w = *static_cast<World *>(h);
    Box huge;
    huge.lo = vec3(-std::numeric_limits<float>::max());
    huge.hi = vec3(std::numeric_limits<float>::max());
    std::vector<uint32_t> m, c;
    std::vector<int32_t> st;
    w.tree.query(UINT32_MAX, huge, m, c, st, nullptr);
    *n_mesh = uint32_t(m.size());
    *n_caps = uint32_t(c.size());
    if (mesh_ids) std::copy(m.begin(), m.end(), mesh_ids);
    if (caps_ids) std::copy(c.begin(), c.end(), caps_ids);
}

void l1_query(void * h, uint32_t caller_capsule, const float * aabb6, uint32_t * mesh_ids, uint32_t * n_mesh,
              uint32_t * caps_ids, uint32_t * n_caps, uint32_t cap) {
    World & w = *static_cast<World *>(h);
    Box box;
    box.lo = vec3(aabb6[0], aabb6[1], aabb6[2]);
    box.hi = vec3(aabb6[3], aabb6[4], aabb6[5]);
    std::vector<uint32_t> m, c;
    std::vector<int32_t> st;
    w.tree.query(caller_capsule, box, m, c, st, nullptr);
    *n_mesh = uint32_t(m.size());
    *n_caps = uint32_t(c.size());
    for (size_t i = 0; i < m.size() && i < cap; ++i) mesh_ids[i] = m[i];
    for (size_t i = 0; i < c.size() && i < cap; ++i) caps_ids[i] = c[i];
}

uint32_t l1_world_cache(void * h, float * xyz, uint32_t cap_floats) {
    World & w = *static_cast<World *>(h);
    uint32_t n = uint32_t(w.cache.size());
    for (uint32_t i = 0; i < n && 3 * i + 2 < cap_floats; ++i) { xyz[3 * i] = w.cache[i].x; xyz[3 * i + 1] = w.cache[i].y; xyz[3 * i + 2] = w.cache[i].z; }
    return n;
}

// physics_system.cpp:205-242
bool l1_raycast(void * h, const float * origin3, const float * dir3, float maxDistance, float * hit3) {
    World & w = *static_cast<World *>(h);
    vec3 origin(origin3[0], origin3[1], origin3[2]);
    vec3 direction(dir3[0], dir3[1], dir3[2]);
    std::vector<uint32_t> tags;
    w.tree.queryRaycast(origin, direction, maxDistance, tags);
    float intersectionTime = std::numeric_limits<float>::max();
    bool intersect = false;
    for (uint32_t id : tags) {
        MeshCol const & m = w.meshes[id];
        uint32_t index = m.first;
        for (uint32_t i = 0; i < m.count; i += 3) {
            float t;
            bool in = intersectLineSegmentTriangle(origin, origin + direction * maxDistance,
                                                   w.cache[index], w.cache[index + 1], w.cache[index + 2], t);
            intersect |= in;
            if (in) intersectionTime = std::min(intersectionTime, t);
            index += 3;
        }
    }
    if (intersect) {
        vec3 hit = origin + direction * intersectionTime * maxDistance;
        hit3[0] = hit.x; hit3[1] = hit.y; hit3[2] = hit.z;
    }
    return intersect;
}

uint32_t l1_trace_num_queries(void * h) { return uint32_t(static_cast<World *>(h)->trace.q_char.size()); }
uint32_t l1_trace_num_hits(void * h) { return uint32_t(static_cast<World *>(h)->trace.h_query.size()); }
uint32_t l1_trace_num_mesh_ids(void * h) { return uint32_t(static_cast<World *>(h)->trace.mesh_ids.size()); }
uint32_t l1_trace_num_caps_ids(void * h) { return uint32_t(static_cast<World *>(h)->trace.caps_ids.size()); }
void l1_trace_get(void * h, uint32_t * q_char, uint32_t * q_sub, float * q_aabb, uint32_t * q_mesh_off,
                  uint32_t * q_caps_off, uint32_t * mesh_ids, uint32_t * caps_ids, uint32_t * q_tri_tests,
                  uint32_t * q_node_tests, uint32_t * h_query, int32_t * h_poly, float * h_impulse,
                  float * h_normal, float * h_pos) {
    Trace & t = static_cast<World *>(h)->trace;
    auto cp = [](auto * dst, auto const & v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    cp(q_char, t.q_char); cp(q_sub, t.q_sub); cp(q_aabb, t.q_aabb); cp(q_mesh_off, t.q_mesh_off);
    cp(q_caps_off, t.q_caps_off); cp(mesh_ids, t.mesh_ids); cp(caps_ids, t.caps_ids);
    cp(q_tri_tests, t.q_tri_tests); cp(q_node_tests, t.q_node_tests); cp(h_query, t.h_query);
    cp(h_poly, t.h_poly); cp(h_impulse, t.h_impulse); cp(h_normal, t.h_normal); cp(h_pos, t.h_pos);
}

} // extern "C"

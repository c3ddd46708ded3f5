// sd_wide_proto.cpp -- CPU prototype used to size the wide-node SignedDistance traversal (tools/, not product).
// Reads the arrays dumped by tools/sd_wide_proto.py (C2 icosphere + the reference-order binary BVH), collapses
// L binary levels into 2^L-wide nodes with oriented child bounds, and counts node visits / leaf tests per query
// for several ways of seeding the prune radius (hint of a Morton neighbour, greedy descent, both, perfect).
//   g++ -O2 -fopenmp -o /tmp/proto/proto tools/sd_wide_proto.cpp && /tmp/proto/proto /tmp/proto
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

struct V3
{
  double x, y, z;
};
static inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 mul(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 u, V3 v) { return {u.y * v.z - v.y * u.z, v.x * u.z - u.x * v.z, u.x * v.y - v.x * u.y}; }

template <typename T>
static std::vector<T> slurp(const std::string& p)
{
  FILE* f = fopen(p.c_str(), "rb");
  if(!f)
  {
    perror(p.c_str());
    exit(1);
  }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<T> v(n / sizeof(T));
  if(fread(v.data(), 1, n, f) != (size_t)n) exit(2);
  fclose(f);
  return v;
}

// Ericson closest point on triangle, squared distance only (plain double; the prototype only counts work)
static double tri_sqdist(V3 p, V3 a, V3 b, V3 c, V3* cpo = nullptr)
{
  V3 ab = sub(b, a), ac = sub(c, a), ap = sub(p, a);
  double d1 = dot(ab, ap), d2 = dot(ac, ap);
  V3 r;
  if(d1 <= 0 && d2 <= 0)
    r = a;
  else
  {
    V3 bp = sub(p, b);
    double d3 = dot(ab, bp), d4 = dot(ac, bp);
    if(d3 >= 0 && d4 <= d3)
      r = b;
    else
    {
      double vc = d1 * d4 - d3 * d2;
      if(vc <= 0 && d1 >= 0 && d3 <= 0)
        r = add(a, mul(ab, d1 / (d1 - d3)));
      else
      {
        V3 cp = sub(p, c);
        double d5 = dot(ab, cp), d6 = dot(ac, cp);
        if(d6 >= 0 && d5 <= d6)
          r = c;
        else
        {
          double vb = d5 * d2 - d1 * d6;
          if(vb <= 0 && d2 >= 0 && d6 <= 0)
            r = add(a, mul(ac, d2 / (d2 - d6)));
          else
          {
            double va = d3 * d6 - d5 * d4;
            if(va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0)
              r = add(b, mul(sub(c, b), (d4 - d3) / ((d4 - d3) + (d5 - d6))));
            else
            {
              double den = 1.0 / (va + vb + vc);
              r = add(a, add(mul(ab, vb * den), mul(ac, vc * den)));
            }
          }
        }
      }
    }
  }
  if(cpo) *cpo = r;
  V3 d = sub(r, p);
  return dot(d, d);
}

struct Obb
{
  float n[3], c[3], h[3];
};

static int N, INNER;
static std::vector<double> verts, inner_nodes;
static std::vector<int32_t> conn, leafs, children;
static std::vector<V3> soup;  // 3 per sorted leaf
static std::vector<int> efirst, elast;
static std::vector<V3> enormal;

static inline int ent_of_child(int c) { return c >= 0 ? c / 2 : INNER + (-c - 1); }

static void frame(const float* nf, V3* A)
{
  V3 n = {nf[0], nf[1], nf[2]};
  double ax = fabs(n.x), ay = fabs(n.y), az = fabs(n.z);
  V3 ek = (ax <= ay && ax <= az) ? V3 {1, 0, 0} : ((ay <= az) ? V3 {0, 1, 0} : V3 {0, 0, 1});
  V3 t1 = cross(n, ek);
  t1 = mul(t1, 1.0 / sqrt(dot(t1, t1)));
  A[0] = n;
  A[1] = t1;
  A[2] = cross(n, t1);
}

static int OBB_MAX = 262144;

static Obb make_obb(int e, const double* org, const double* aabb /*6*/)
{
  Obb o;
  int first = efirst[e], last = elast[e];
  if(last - first + 1 > OBB_MAX)
  {
    o.n[0] = 0;
    o.n[1] = 0;
    o.n[2] = 1;  // frame() of (0,0,1): t1 = n x e_x = (0,1,0)... fine, axis aligned
    V3 A[3];
    frame(o.n, A);
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for(int cx = 0; cx < 8; ++cx)
    {
      V3 p = {aabb[(cx & 1) ? 3 : 0] - org[0], aabb[(cx & 2) ? 4 : 1] - org[1], aabb[(cx & 4) ? 5 : 2] - org[2]};
      for(int k = 0; k < 3; ++k)
      {
        double d = dot(A[k], p);
        lo[k] = std::min(lo[k], d);
        hi[k] = std::max(hi[k], d);
      }
    }
    for(int k = 0; k < 3; ++k)
    {
      o.c[k] = (float)(0.5 * (lo[k] + hi[k]));
      o.h[k] = (float)(std::max(hi[k] - o.c[k], o.c[k] - lo[k]) * (1 + 1e-6));
    }
    return o;
  }
  V3 ns = enormal[e];
  double len = sqrt(dot(ns, ns));
  V3 n = len > 1e-140 ? mul(ns, 1.0 / len) : V3 {0, 0, 1};
  o.n[0] = (float)n.x;
  o.n[1] = (float)n.y;
  o.n[2] = (float)n.z;
  V3 A[3];
  frame(o.n, A);
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for(int p = first; p <= last; ++p)
    for(int j = 0; j < 3; ++j)
    {
      V3 v = soup[3 * (size_t)p + j];
      V3 r = {v.x - org[0], v.y - org[1], v.z - org[2]};
      for(int k = 0; k < 3; ++k)
      {
        double d = dot(A[k], r);
        lo[k] = std::min(lo[k], d);
        hi[k] = std::max(hi[k], d);
      }
    }
  for(int k = 0; k < 3; ++k)
  {
    o.c[k] = (float)(0.5 * (lo[k] + hi[k]));
    double hh = std::max(hi[k] - (double)o.c[k], (double)o.c[k] - lo[k]);
    o.h[k] = nextafterf((float)(hh * (1 + 1e-7) + 1e-30), INFINITY);
  }
  return o;
}

struct Wide
{
  int child[8];  // >= 0 wide index, < 0 leaf -(pos+1), INT32_MIN empty
  double org[3];
  Obb ob[8];
  int bnode[7];  // binary node ids: [0] root, [1..2] level 2, [3..6] level 3 (-1 = absent)
  int nslots;
};
static std::vector<Wide> wide;

static inline double obb_lb2(const Obb& o, const double* r, double slop)
{
  V3 A[3];
  frame(o.n, A);
  double s = 0;
  V3 rv = {r[0], r[1], r[2]};
  for(int k = 0; k < 3; ++k)
  {
    double d = dot(A[k], rv);
    double t = fabs(d - o.c[k]) - o.h[k] - slop;
    if(t > 0) s += t * t;
  }
  return s * (1.0 - 1e-6);
}

static inline double thr_of(double sq)
{
  if(sq >= 1e300) return DBL_MAX;
  double d = sqrt(sq) + 1.0000001e-6;
  return d * d * (1 + 1e-12);
}

static int LEVELS = 3;

static void build_wide()
{
  const int W = 1 << LEVELS;
  wide.clear();
  std::vector<int> broot;  // binary root of each wide node
  broot.push_back(0);
  wide.push_back(Wide());
  for(size_t w = 0; w < broot.size(); ++w)
  {
    int b = broot[w];
    Wide wd;
    for(int s = 0; s < 8; ++s) wd.child[s] = INT32_MIN;
    for(int k = 0; k < 7; ++k) wd.bnode[k] = -1;
    // org = centroid of b's box (union of both children)
    double lo[3], hi[3];
    for(int d = 0; d < 3; ++d)
    {
      lo[d] = std::min(inner_nodes[(2 * (size_t)b) * 6 + d], inner_nodes[(2 * (size_t)b + 1) * 6 + d]);
      hi[d] = std::max(inner_nodes[(2 * (size_t)b) * 6 + 3 + d], inner_nodes[(2 * (size_t)b + 1) * 6 + 3 + d]);
      wd.org[d] = 0.5 * (lo[d] + hi[d]);
    }
    // expand
    struct It
    {
      int child;  // binary child code
      int slot;   // slot prefix
      int level;
      int bslot;  // index in bnode numbering of this node if inner
      const double* box;
    };
    std::vector<It> st;
    wd.bnode[0] = b;
    wd.nslots = 0;
    std::vector<It> cur;
    cur.push_back({2 * b, 0, 0, 0, nullptr});
    // recursive expansion
    std::vector<It> work;
    work.push_back({2 * b, 0, 0, 0, nullptr});
    while(!work.empty())
    {
      It it = work.back();
      work.pop_back();
      if(it.child >= 0 && it.level < LEVELS)
      {
        int bi = it.child / 2;
        if(it.level > 0) wd.bnode[it.bslot] = bi;
        for(int s = 0; s < 2; ++s)
        {
          It ch;
          ch.child = children[2 * (size_t)bi + s];
          ch.level = it.level + 1;
          ch.slot = it.slot | (s << (LEVELS - 1 - it.level));
          ch.bslot = it.level == 0 ? 1 + s : (it.level == 1 ? 3 + 2 * (it.bslot - 1) + s : -1);
          ch.box = &inner_nodes[(2 * (size_t)bi + s) * 6];
          work.push_back(ch);
        }
      }
      else
      {
        // terminal: leaf, or inner node at the bottom level
        int e = ent_of_child(it.child);
        bool valid = !(it.box[0] > it.box[3]);
        if(!valid) continue;
        wd.ob[it.slot] = make_obb(e, wd.org, it.box);
        if(it.child < 0)
          wd.child[it.slot] = it.child;
        else
        {
          wd.child[it.slot] = (int)broot.size();
          broot.push_back(it.child / 2);
          wide.push_back(Wide());
        }
        wd.nslots++;
      }
    }
    wide[w] = wd;
    (void)W;
  }
}

struct Stats
{
  double wide_visits = 0, slot_tests = 0, leaf_tests = 0, greedy_visits = 0, multi = 0, ambiguous = 0;
  long maxv = 0;
  std::vector<int> hist;
};

// swap bit of binary node b for query q (reference: right first iff dl > dr)
static inline bool swap_of(int b, const double* q)
{
  const double* L = &inner_nodes[(2 * (size_t)b) * 6];
  const double* R = &inner_nodes[(2 * (size_t)b + 1) * 6];
  double dl = 0, dr = 0;
  for(int d = 0; d < 3; ++d)
  {
    double c = 0.5 * (L[d] + L[3 + d]) - q[d];
    dl += c * c;
  }
  if(R[0] > R[3])
    dr = DBL_MAX;
  else
    for(int d = 0; d < 3; ++d)
    {
      double c = 0.5 * (R[d] + R[3 + d]) - q[d];
      dr += c * c;
    }
  return dl > dr;
}

struct QRes
{
  double sq;
  int pos;
  V3 cp;
};

// mode bits: 1 = use hint, 2 = greedy descent, 4 = perfect (true minimum given), 8 = deferred leaf evaluation (thr fixed during DFS)
static QRes run_query(const double* q, int mode, const QRes* hint, double perfect_sq, Stats& S, double slop)
{
  const V3 qv = {q[0], q[1], q[2]};
  double ub = DBL_MAX;
  QRes best {DBL_MAX, -1, {0, 0, 0}};
  if((mode & 1) && hint && hint->pos >= 0)
  {
    double s = tri_sqdist(qv, soup[3 * (size_t)hint->pos], soup[3 * (size_t)hint->pos + 1], soup[3 * (size_t)hint->pos + 2]);
    ub = std::min(ub, s);
    S.leaf_tests += 1;
  }
  if(mode & 2)
  {
    int w = 0;
    while(true)
    {
      const Wide& wd = wide[w];
      S.greedy_visits += 1;
      double r[3] = {q[0] - wd.org[0], q[1] - wd.org[1], q[2] - wd.org[2]};
      double bl = DBL_MAX;
      int bs = -1;
      for(int s = 0; s < 8; ++s)
        if(wd.child[s] != INT32_MIN)
        {
          double l = obb_lb2(wd.ob[s], r, slop);
          if(l < bl)
          {
            bl = l;
            bs = s;
          }
        }
      int c = wd.child[bs];
      if(c < 0)
      {
        int pos = -c - 1;
        double s = tri_sqdist(qv, soup[3 * (size_t)pos], soup[3 * (size_t)pos + 1], soup[3 * (size_t)pos + 2]);
        S.leaf_tests += 1;
        ub = std::min(ub, s);
        break;
      }
      w = c;
    }
  }
  if(mode & 4) ub = perfect_sq;
  double thr = thr_of(ub);
  // ordered DFS
  int stack[512];
  int sp = 0;
  stack[sp++] = 0;
  long visits = 0;
  while(sp > 0)
  {
    int c = stack[--sp];
    if(c < 0)
    {
      int pos = -c - 1;
      V3 cp;
      double s = tri_sqdist(qv, soup[3 * (size_t)pos], soup[3 * (size_t)pos + 1], soup[3 * (size_t)pos + 2], &cp);
      S.leaf_tests += 1;
      if(s < best.sq)
      {
        best.sq = s;
        best.pos = pos;
        best.cp = cp;
      }
      if(!(mode & 8)) thr = std::min(thr, thr_of(best.sq));
      continue;
    }
    const Wide& wd = wide[c];
    ++visits;
    double r[3] = {q[0] - wd.org[0], q[1] - wd.org[1], q[2] - wd.org[2]};
    unsigned mask = 0;
    int cnt = 0;
    for(int s = 0; s < 8; ++s)
      if(wd.child[s] != INT32_MIN)
      {
        S.slot_tests += 1;
        if(obb_lb2(wd.ob[s], r, slop) <= thr)
        {
          mask |= 1u << s;
          ++cnt;
        }
      }
    if(cnt == 0) continue;
    if(cnt > 1) S.multi += 1;
    // order: key bits from swap bits
    bool sw[7];
    for(int k = 0; k < 7; ++k) sw[k] = (cnt > 1 && wd.bnode[k] >= 0) ? swap_of(wd.bnode[k], q) : false;
    for(int key = (1 << LEVELS) - 1; key >= 0; --key)
    {
      int slot = 0;
      if(LEVELS == 3)
      {
        int k1 = (key >> 2) & 1, k2 = (key >> 1) & 1, k3 = key & 1;
        int b1 = k1 ^ (int)sw[0];
        int b2 = k2 ^ (int)sw[1 + b1];
        int b3 = k3 ^ (int)sw[3 + 2 * b1 + b2];
        slot = (b1 << 2) | (b2 << 1) | b3;
      }
      else if(LEVELS == 2)
      {
        int k1 = (key >> 1) & 1, k2 = key & 1;
        int b1 = k1 ^ (int)sw[0];
        int b2 = k2 ^ (int)sw[1 + b1];
        slot = (b1 << 1) | b2;
      }
      else
      {
        slot = key ^ (int)sw[0];
      }
      if(mask & (1u << slot))
      {
        if(sp >= 511)
        {
          fprintf(stderr, "stack overflow\n");
          exit(3);
        }
        stack[sp++] = wd.child[slot];
      }
    }
  }
  S.wide_visits += visits;
  S.maxv = std::max(S.maxv, visits);
  int hb = visits >= 1024 ? 11 : (int)floor(log2((double)std::max(1L, visits)));
  if((int)S.hist.size() < 12) S.hist.resize(12, 0);
  S.hist[hb]++;
  return best;
}


// ---- packet simulation: 32 Morton-consecutive queries share the upper part of the walk (binary nodes, LEVELS must be 1) ----
struct PStats
{
  double packet_leaf = 0, packet_nodes = 0, packet_lane_tests = 0, private_visits = 0, leaf_tests = 0, private_pushes = 0, maxpriv = 0, queries = 0;
};
static void run_packet(const double* Q /*32 x 3*/, const QRes* hints /*32 or null*/, QRes* out, int T, bool recheck, PStats& S)
{
  double thr[32], best[32];
  int bpos[32];
  for(int l = 0; l < 32; ++l)
  {
    best[l] = DBL_MAX;
    bpos[l] = -1;
    thr[l] = DBL_MAX;
    if(hints && hints[l].pos >= 0)
    {
      V3 qv = {Q[3 * l], Q[3 * l + 1], Q[3 * l + 2]};
      double s = tri_sqdist(qv, soup[3 * (size_t)hints[l].pos], soup[3 * (size_t)hints[l].pos + 1], soup[3 * (size_t)hints[l].pos + 2]);
      thr[l] = thr_of(s);
      S.leaf_tests += 1;
    }
  }
  auto leaf = [&](int l, int pos) {
    V3 qv = {Q[3 * l], Q[3 * l + 1], Q[3 * l + 2]};
    double s = tri_sqdist(qv, soup[3 * (size_t)pos], soup[3 * (size_t)pos + 1], soup[3 * (size_t)pos + 2]);
    S.leaf_tests += 1;
    if(s < best[l])
    {
      best[l] = s;
      bpos[l] = pos;
    }
    thr[l] = std::min(thr[l], thr_of(best[l]));
  };
  struct PE
  {
    int node;
    unsigned mask;
  };
  const bool defer = getenv("DEFER_LEAF") && hints;
  if(!hints) T = 33;
  std::vector<PE> ws;
  struct LE
  {
    int node;
    double lb;
  };
  std::vector<LE> priv[32];
  ws.push_back({0, 0xffffffffu});
  while(!ws.empty())
  {
    PE e = ws.back();
    ws.pop_back();
    const Wide& wd = wide[e.node];
    S.packet_nodes += 1;
    unsigned m[2] = {0, 0};
    double lbs[2][32];
    for(int l = 0; l < 32; ++l)
      if(e.mask >> l & 1)
      {
        S.packet_lane_tests += 1;
        double r[3] = {Q[3 * l] - wd.org[0], Q[3 * l + 1] - wd.org[1], Q[3 * l + 2] - wd.org[2]};
        for(int s = 0; s < 2; ++s)
          if(wd.child[s] != INT32_MIN)
          {
            lbs[s][l] = obb_lb2(wd.ob[s], r, 0.0);
            if(lbs[s][l] <= thr[l]) m[s] |= 1u << l;
          }
      }
    // visit the child most lanes want last-pushed (first popped)
    int order[2] = {0, 1};
    if(__builtin_popcount(m[0]) > __builtin_popcount(m[1])) std::swap(order[0], order[1]);
    for(int k = 0; k < 2; ++k)
    {
      int s = order[k];
      if(!m[s]) continue;
      int c = wd.child[s];
      if(c < 0 && !defer)
      {
        for(int l = 0; l < 32; ++l)
          if(m[s] >> l & 1)
            if(lbs[s][l] <= thr[l]) { leaf(l, -c - 1); S.packet_leaf += 1; }
      }
      else if(c >= 0 && __builtin_popcount(m[s]) >= T)
        ws.push_back({c, m[s]});
      else
        for(int l = 0; l < 32; ++l)
          if(m[s] >> l & 1)
          {
            priv[l].push_back({c, lbs[s][l]});
            S.private_pushes += 1;
          }
    }
  }
  for(int l = 0; l < 32; ++l)
  {
    S.maxpriv = std::max(S.maxpriv, (double)priv[l].size());
    std::vector<LE>& st = priv[l];
    if(getenv("BEST_FIRST") && st.size() > 1)
    {
      if(atoi(getenv("BEST_FIRST")) == 2)
        std::sort(st.begin(), st.end(), [](const LE& a, const LE& b) { return a.lb > b.lb; });
      else
      {
        size_t bi = 0;
        for(size_t k = 1; k < st.size(); ++k)
          if(st[k].lb < st[bi].lb) bi = k;
        std::swap(st[bi], st.back());
      }
    }
    while(!st.empty())
    {
      LE e = st.back();
      st.pop_back();
      if(recheck && e.lb > thr[l]) continue;
      if(e.node < 0)
      {
        leaf(l, -e.node - 1);
        continue;
      }
      const Wide& wd = wide[e.node];
      S.private_visits += 1;
      double r[3] = {Q[3 * l] - wd.org[0], Q[3 * l + 1] - wd.org[1], Q[3 * l + 2] - wd.org[2]};
      double lb[2] = {DBL_MAX, DBL_MAX};
      for(int s = 0; s < 2; ++s)
        if(wd.child[s] != INT32_MIN) lb[s] = obb_lb2(wd.ob[s], r, 0.0);
      int first = lb[0] <= lb[1] ? 0 : 1;  // nearer first: push the farther one
      int second = 1 - first;
      if(lb[second] <= thr[l]) st.push_back({wd.child[second], lb[second]});
      if(lb[first] <= thr[l]) st.push_back({wd.child[first], lb[first]});
    }
    out[l].sq = best[l];
    out[l].pos = bpos[l];
    S.queries += 1;
  }
}

static inline uint32_t compact3(uint32_t v)
{
  v &= 0x09249249;
  v = (v ^ (v >> 2)) & 0x030c30c3;
  v = (v ^ (v >> 4)) & 0x0300f00f;
  v = (v ^ (v >> 8)) & 0xff0000ff;
  v = (v ^ (v >> 16)) & 0x000003ff;
  return v;
}

int main(int argc, char** argv)
{
  std::string dir = argc > 1 ? argv[1] : "/tmp/proto";
  if(argc > 2) LEVELS = atoi(argv[2]);
  int hint_dist = argc > 3 ? atoi(argv[3]) : 32;
  double slop_rel = argc > 4 ? atof(argv[4]) : 0.0;
  verts = slurp<double>(dir + "/verts.bin");
  conn = slurp<int32_t>(dir + "/conn.bin");
  leafs = slurp<int32_t>(dir + "/leafs.bin");
  children = slurp<int32_t>(dir + "/children.bin");
  inner_nodes = slurp<double>(dir + "/inner_nodes.bin");
  N = (int)leafs.size();
  INNER = N - 1;
  soup.resize(3 * (size_t)N);
  for(int p = 0; p < N; ++p)
    for(int j = 0; j < 3; ++j)
    {
      int v = conn[3 * (size_t)leafs[p] + j];
      soup[3 * (size_t)p + j] = {verts[3 * (size_t)v], verts[3 * (size_t)v + 1], verts[3 * (size_t)v + 2]};
    }
  // ranges + normal sums: post-order
  efirst.assign(INNER + N, 0);
  elast.assign(INNER + N, 0);
  enormal.assign(INNER + N, V3 {0, 0, 0});
  for(int p = 0; p < N; ++p)
  {
    efirst[INNER + p] = elast[INNER + p] = p;
    enormal[INNER + p] = cross(sub(soup[3 * (size_t)p + 1], soup[3 * (size_t)p]), sub(soup[3 * (size_t)p + 2], soup[3 * (size_t)p]));
  }
  {
    std::vector<int> order;
    order.reserve(INNER);
    std::vector<int> st {0};
    while(!st.empty())
    {
      int b = st.back();
      st.pop_back();
      order.push_back(b);
      for(int s = 0; s < 2; ++s)
      {
        int c = children[2 * (size_t)b + s];
        if(c >= 0) st.push_back(c / 2);
      }
    }
    for(int i = (int)order.size() - 1; i >= 0; --i)
    {
      int b = order[i];
      int e0 = ent_of_child(children[2 * (size_t)b]), e1 = ent_of_child(children[2 * (size_t)b + 1]);
      efirst[b] = std::min(efirst[e0], efirst[e1]);
      elast[b] = std::max(elast[e0], elast[e1]);
      enormal[b] = add(enormal[e0], enormal[e1]);
    }
  }
  build_wide();
  {
    double fill = 0;
    for(auto& w : wide) fill += w.nslots;
    printf("LEVELS %d: %zu wide nodes for %d leaves (%.2f slots used of %d), %.1f MB at 384 B\n", LEVELS, wide.size(), N, fill / wide.size(),
           1 << LEVELS, wide.size() * 384.0 / 1e6);
  }
  // queries: runs of consecutive Morton indices of the 256^3 grid
  const int RUNS = 96, RUNLEN = 2048;
  std::vector<double> Q;
  for(int r = 0; r < RUNS; ++r)
  {
    uint32_t m0 = (uint32_t)((double)r / RUNS * (1u << 24));
    for(int i = 0; i < RUNLEN; ++i)
    {
      uint32_t m = m0 + i;
      uint32_t ix = compact3(m), iy = compact3(m >> 1), iz = compact3(m >> 2);
      Q.push_back(-1.0 + ix * (2.0 / 255));
      Q.push_back(-1.0 + iy * (2.0 / 255));
      Q.push_back(-1.0 + iz * (2.0 / 255));
    }
  }
  const int nq = RUNS * RUNLEN;
  std::vector<QRes> exact(nq);
  const char* names[] = {"greedy", "perfect", "hint", "greedy+hint", "greedy deferred", "greedy+hint deferred", "none"};
  const int modes[] = {2, 4, 1, 3, 2 | 8, 3 | 8, 0};
  for(int v = 0; v < 7; ++v)
  {
    if(v == 6 && LEVELS != 1 && argc < 6) continue;
    Stats tot;
    tot.hist.assign(12, 0);
#pragma omp parallel
    {
      Stats S;
      S.hist.assign(12, 0);
#pragma omp for schedule(dynamic, 1)
      for(int r = 0; r < RUNS; ++r)
      {
        std::vector<QRes> res(RUNLEN);
        for(int i = 0; i < RUNLEN; ++i)
        {
          const double* q = &Q[3 * ((size_t)r * RUNLEN + i)];
          double slop = slop_rel * (fabs(q[0]) + fabs(q[1]) + fabs(q[2]) + 1.0);
          const QRes* hint = i >= hint_dist ? &res[i - hint_dist] : nullptr;
          double psq = v == 0 ? 0 : 0;
          if(modes[v] & 4) psq = exact[(size_t)r * RUNLEN + i].sq;
          res[i] = run_query(q, modes[v], hint, psq, S, slop);
          if(v != 0 && res[i].sq != exact[(size_t)r * RUNLEN + i].sq)
          {
            fprintf(stderr, "MISMATCH variant %d run %d i %d: %.17g vs %.17g\n", v, r, i, res[i].sq, exact[(size_t)r * RUNLEN + i].sq);
          }
        }
        if(v == 0)
          for(int i = 0; i < RUNLEN; ++i) exact[(size_t)r * RUNLEN + i] = res[i];
      }
#pragma omp critical
      {
        tot.wide_visits += S.wide_visits;
        tot.slot_tests += S.slot_tests;
        tot.leaf_tests += S.leaf_tests;
        tot.greedy_visits += S.greedy_visits;
        tot.multi += S.multi;
        tot.maxv = std::max(tot.maxv, S.maxv);
        for(int k = 0; k < 12; ++k) tot.hist[k] += S.hist[k];
      }
    }
    if(v == 0)
    {
      // the first pass ran "perfect" with psq = 0 -> thr tiny -> wrong; redo properly: compute exact by unbounded DFS
    }
    printf("%-22s dfs visits %.1f (max %ld) greedy %.1f slot tests %.1f leaf tests %.2f multi-nodes %.1f | hist", names[v], tot.wide_visits / nq,
           tot.maxv, tot.greedy_visits / nq, tot.slot_tests / nq, tot.leaf_tests / nq, tot.multi / nq);
    for(int k = 0; k < 12; ++k) printf(" %d", tot.hist[k]);
    printf("\n");
  }
  if(LEVELS == 1)
  {
    for(int T : {8, 12, 16, 20, 24})
      for(int recheck = 1; recheck < 2; ++recheck)
      {
        PStats tot;
        long mism = 0;
#pragma omp parallel
        {
          PStats S;
          long mm = 0;
#pragma omp for schedule(dynamic, 1)
          for(int r = 0; r < RUNS; ++r)
          {
            std::vector<QRes> res(RUNLEN);
            for(int p = 0; p < RUNLEN / 32; ++p)
            {
              const double* q = &Q[3 * ((size_t)r * RUNLEN + 32 * p)];
              run_packet(q, p ? &res[32 * (p - 1)] : nullptr, &res[32 * p], T, recheck != 0, S);
              for(int l = 0; l < 32; ++l)
                if(res[32 * p + l].sq != exact[(size_t)r * RUNLEN + 32 * p + l].sq) ++mm;
            }
          }
#pragma omp critical
          {
            tot.packet_nodes += S.packet_nodes; tot.packet_leaf += S.packet_leaf;
            tot.packet_lane_tests += S.packet_lane_tests;
            tot.private_visits += S.private_visits;
            tot.leaf_tests += S.leaf_tests;
            tot.private_pushes += S.private_pushes;
            tot.maxpriv = std::max(tot.maxpriv, S.maxpriv);
            tot.queries += S.queries;
            mism += mm;
          }
        }
        printf("packet T=%2d recheck=%d: packet nodes/packet %.1f (lanes active %.1f), private visits/query %.1f, pushes/query %.1f (max %g), leaf tests/query %.2f (in packet phase %.2f), mismatches %ld\n",
               T, recheck, tot.packet_nodes / (tot.queries / 32), tot.packet_lane_tests / std::max(1.0, tot.packet_nodes), tot.private_visits / tot.queries,
               tot.private_pushes / tot.queries, tot.maxpriv, tot.leaf_tests / tot.queries, tot.packet_leaf / tot.queries, mism);
      }
  }
  return 0;
}

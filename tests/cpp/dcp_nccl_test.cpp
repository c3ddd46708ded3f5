// dcp_nccl_test.cpp -- the distributed cases from plain host C++, no launcher: the parent forks one process per rank
// (one GPU each), rank 0 passes the NCCL id to the others through a file, and every rank drives the C++ shims
//   axom_b200::quest::Communicator / DistributedClosestPoint::computeClosestPoints   (axb_dcp_compute_closest_points)
//   axb_sd_compute_distances_minreduce                                               (BASELINE config C5)
// exactly as a host code would use the reference's classes with an MPI communicator
// (quest/DistributedClosestPoint.hpp:95-166, quest/tests/quest_distributed_distance_query_example.cpp).
//
// Checked on every rank against a brute-force search over ALL ranks' object points with the reference's tie rule: the
// smallest squared distance (same expression, separately rounded), among equal ones the first rank in ring order from
// the query's owner (DistributedClosestPointImpl.hpp:737-880).  Object points are exact duplicates ACROSS ranks in
// places, and some queries sit exactly on them, so the rule is exercised; inside a rank distances are distinct.
//
//   dcp_nccl_test <nranks> [scratch-dir]      (nranks <= number of GPUs; 1 runs the whole protocol on one rank)
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <string>
#include <vector>

#include "axom_b200/DistributedClosestPoint.hpp"

namespace ab = axom_b200;

struct Lcg
{
  unsigned long long s;
  explicit Lcg(unsigned long long seed) : s(seed * 2862933555777941757ull + 3037000493ull) { }
  double next()  // uniform in [0, 1), 53 bits
  {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(s >> 11) * (1.0 / 9007199254740992.0);
  }
};

template <int D>
struct Case
{
  std::vector<std::vector<double>> obj;    // per rank, interleaved
  std::vector<std::vector<int>> dom;       // per rank, domain id per point
  std::vector<std::vector<double>> query;  // per rank, interleaved
};

template <int D>
static Case<D> make_case(int N, int variant)
{
  Case<D> c;
  c.obj.resize(N);
  c.dom.resize(N);
  c.query.resize(N);
  Lcg rng(1234 + 17 * variant + D);
  for(int r = 0; r < N; ++r)
  {
    int n = 4000 + 700 * r;
    if(variant == 1 && r == 1 % N && N > 1) n = 0;  // a rank without object points
    if(variant == 2 && r == N - 1) n = 1;           // a rank with a single point
    for(int i = 0; i < n; ++i)
    {
      // every rank covers the unit cube, shifted a little: boxes overlap, so the nearest-centre rule spreads the work
      for(int d = 0; d < D; ++d) c.obj[r].push_back(rng.next() + 0.15 * r * (d == 0));
      c.dom[r].push_back(10 * r + (i < n / 2 ? 0 : 1));
    }
  }
  // exact duplicates across ranks: the first 40 points of rank 0 also exist on every other (non-empty) rank
  for(int r = 1; r < N; ++r)
    for(int i = 0; i < 40 && (size_t)(i * D + D) <= c.obj[r].size() && (size_t)(i * D + D) <= c.obj[0].size(); ++i)
      for(int d = 0; d < D; ++d) c.obj[r][i * D + d] = c.obj[0][i * D + d];
  for(int r = 0; r < N; ++r)
  {
    int nq = 3000 + 500 * r;
    if(variant == 1 && r == 0 && N > 1) nq = 0;  // a rank without queries
    for(int i = 0; i < nq; ++i)
      for(int d = 0; d < D; ++d) c.query[r].push_back(rng.next() * 1.6 - 0.2);
    // queries exactly on duplicated points: ties at distance zero, decided by ring order
    for(int i = 0; i < 30 && i < nq && (size_t)(i * D + D) <= c.obj[0].size(); ++i)
      for(int d = 0; d < D; ++d) c.query[r][i * D + d] = c.obj[0][i * D + d];
  }
  return c;
}

template <int D>
static int check_case(ab::quest::Communicator* comm, int N, int rank, int variant, double threshold)
{
  using DCP = ab::quest::DistributedClosestPoint<D>;
  using Pt = typename DCP::PointType;
  const Case<D> c = make_case<D>(N, variant);
  DCP dcp(/*device*/ rank, rank);
  if(comm) dcp.setCommunicator(*comm);
  if(threshold >= 0.0) dcp.setDistanceThreshold(threshold);
  {
    // two domains per rank, as importObjectPoints flattens them
    std::vector<std::vector<Pt>> doms(2);
    std::vector<ab::IndexType> ids = {10 * rank, 10 * rank + 1};
    const size_t n = c.dom[rank].size();
    for(size_t i = 0; i < n; ++i)
    {
      Pt p;
      for(int d = 0; d < D; ++d) p[d] = c.obj[rank][i * D + d];
      doms[c.dom[rank][i] == 10 * rank ? 0 : 1].push_back(p);
    }
    dcp.setObjectMesh(doms, ids);
  }
  dcp.generateBVHTree();
  const int nq = (int)(c.query[rank].size() / D);
  std::vector<ab::IndexType> cp_index(nq), cp_dom(nq), cp_rank(nq);
  std::vector<Pt> cp_coords(nq);
  std::vector<double> cp_dist(nq);
  dcp.computeClosestPoints(reinterpret_cast<const Pt*>(c.query[rank].data()), nq, cp_index.data(), cp_dom.data(), cp_rank.data(), cp_coords.data(),
                           cp_dist.data());
  // brute force with the reference's tie rule
  const double sq_th = threshold >= 0.0 ? threshold * threshold : std::numeric_limits<double>::max();
  int bad = 0, ties = 0, missing = 0;
  for(int i = 0; i < nq; ++i)
  {
    const double* q = &c.query[rank][(size_t)i * D];
    double best = std::numeric_limits<double>::max();
    int brank = -1, bidx = -1;
    for(int k = 0; k < N; ++k)
    {
      const int r = (rank + k) % N;  // ring order from the owner
      const size_t n = c.dom[r].size();
      double rbest = std::numeric_limits<double>::max();
      int ridx = -1;
      for(size_t j = 0; j < n; ++j)
      {
        double s = 0.0;
        for(int d = 0; d < D; ++d)
        {
          const double v = c.obj[r][j * D + d] - q[d];
          s += v * v;
        }
        if(s < rbest)
        {
          rbest = s;
          ridx = (int)j;
        }
      }
      if(ridx >= 0 && rbest <= sq_th)
      {
        if(rbest == best && brank >= 0) ++ties;
        if(rbest < best)
        {
          best = rbest;
          brank = r;
          bidx = ridx;
        }
      }
    }
    bool ok;
    if(brank < 0)
    {
      ++missing;
      ok = cp_rank[i] == -1 && cp_index[i] == -1 && cp_dom[i] == -1 && std::isnan(cp_dist[i]) && std::isnan(cp_coords[i][0]);
    }
    else
    {
      ok = cp_rank[i] == brank && cp_index[i] == bidx && cp_dom[i] == c.dom[brank][bidx] && cp_dist[i] == std::sqrt(best);
      for(int d = 0; d < D; ++d) ok = ok && cp_coords[i][d] == c.obj[brank][(size_t)bidx * D + d];
    }
    if(!ok && bad++ < 5)
      std::fprintf(stderr, "[rank %d] D=%d variant %d query %d: got rank %d index %d dist %.17g, want rank %d index %d dist %.17g\n", rank, D,
                   variant, i, cp_rank[i], cp_index[i], cp_dist[i], brank, bidx, brank >= 0 ? std::sqrt(best) : NAN);
  }
  std::printf("[rank %d] DCP D=%d variant %d threshold %g: %d queries, %d cross-rank ties, %d beyond the threshold, %d wrong\n", rank, D, variant,
              threshold, nq, ties, missing, bad);
  if(N > 1 && variant == 0 && threshold < 0.0 && ties < 20) return 1;  // the tie rule must really be exercised
  return bad == 0 ? 0 : 1;
}

// C5: random triangles dealt round-robin to the ranks; MIN over the parts == the smallest of the parts' own answers
static int check_minreduce(ab::quest::Communicator& comm, int N, int rank)
{
  Lcg rng(77);
  const int ntri = 3000, nq = 20000;
  std::vector<double> tri((size_t)ntri * 9);
  for(int t = 0; t < ntri; ++t)
  {
    double c[3] = {rng.next(), rng.next(), rng.next()};
    for(int v = 0; v < 3; ++v)
      for(int d = 0; d < 3; ++d) tri[(size_t)t * 9 + v * 3 + d] = c[d] + 0.05 * (rng.next() - 0.5);
  }
  std::vector<double> q((size_t)nq * 3);
  for(double& v : q) v = rng.next() * 1.4 - 0.2;
  auto make_part = [&](int r, axb_sd** out) {
    std::vector<double> x, y, z;
    std::vector<int32_t> conn;
    for(int t = r; t < ntri; t += N)
      for(int v = 0; v < 3; ++v)
      {
        conn.push_back((int32_t)x.size());
        x.push_back(tri[(size_t)t * 9 + v * 3 + 0]);
        y.push_back(tri[(size_t)t * 9 + v * 3 + 1]);
        z.push_back(tri[(size_t)t * 9 + v * 3 + 2]);
      }
    return axb_sd_create(out, rank, x.data(), y.data(), z.data(), (int32_t)x.size(), conn.data(), nullptr, (int32_t)(conn.size() / 3), 3, AXB_MEM_HOST,
                         /*watertight*/ 0, /*compute_sign*/ 0);
  };
  axb_array_desc qd;
  std::memset(&qd, 0, sizeof(qd));
  for(int k = 0; k < 3; ++k) qd.comp[k] = q.data() + k;
  qd.stride_bytes = 24;
  qd.ncomp = 3;
  qd.memspace = AXB_MEM_HOST;
  std::vector<double> want(nq, std::numeric_limits<double>::max()), part(nq), got(nq);
  for(int r = 0; r < N; ++r)  // every part evaluated locally, no communication
  {
    axb_sd* s = nullptr;
    if(make_part(r, &s) != AXB_OK || axb_sd_compute_distances(s, &qd, nq, part.data(), nullptr, nullptr, AXB_MEM_HOST) != AXB_OK)
    {
      std::fprintf(stderr, "[rank %d] local part %d failed: %s\n", rank, r, axb_last_error());
      return 1;
    }
    for(int i = 0; i < nq; ++i) want[i] = part[i] < want[i] ? part[i] : want[i];
    axb_sd_destroy(s);
  }
  axb_sd* mine = nullptr;
  if(make_part(rank, &mine) != AXB_OK || axb_sd_compute_distances_minreduce(mine, comm.handle(), &qd, nq, got.data(), AXB_MEM_HOST) != AXB_OK)
  {
    std::fprintf(stderr, "[rank %d] minreduce failed: %s\n", rank, axb_last_error());
    return 1;
  }
  axb_sd_destroy(mine);
  int bad = 0;
  for(int i = 0; i < nq; ++i) bad += std::memcmp(&want[i], &got[i], sizeof(double)) != 0;
  int64_t bytes = 0, calls = 0;
  axb_comm_get_traffic(comm.handle(), &bytes, &calls);
  std::printf("[rank %d] C5 minreduce: %d queries over %d surface parts, %d differ from the MIN of the parts (%lld bytes in %lld collectives so far)\n",
              rank, nq, N, bad, (long long)bytes, (long long)calls);
  return bad == 0 ? 0 : 1;
}

static int run_rank(int N, int rank, const std::string& id_file)
{
  using ab::quest::Communicator;
  Communicator::Id id;
  if(rank == 0)
  {
    id = Communicator::uniqueId();
    const std::string tmp = id_file + ".tmp";
    std::ofstream(tmp, std::ios::binary).write(reinterpret_cast<const char*>(id.data()), (std::streamsize)id.size());
    std::rename(tmp.c_str(), id_file.c_str());
  }
  else
  {
    for(int tries = 0; tries < 12000 && access(id_file.c_str(), R_OK) != 0; ++tries) usleep(10000);
    std::ifstream f(id_file, std::ios::binary);
    id.resize(AXB_COMM_ID_BYTES);
    f.read(reinterpret_cast<char*>(id.data()), (std::streamsize)id.size());
    if(!f) return 3;
  }
  Communicator comm(N, rank, id, /*device*/ rank);
  if(rank == 0) std::printf("NCCL: %s\n", axb_comm_library());
  int fail = 0;
  fail += check_case<3>(&comm, N, rank, 0, -1.0);
  fail += check_case<2>(&comm, N, rank, 0, -1.0);
  fail += check_case<3>(&comm, N, rank, 1, -1.0);   // an empty object rank, a rank without queries
  fail += check_case<3>(&comm, N, rank, 2, 0.03);   // distance threshold: far queries stay unset
  fail += check_case<2>(&comm, N, rank, 2, 0.0);    // threshold 0: only exact hits
  fail += check_minreduce(comm, N, rank);
  if(N == 1) fail += check_case<3>(nullptr, 1, 0, 0, -1.0);  // no communicator at all
  return fail == 0 ? 0 : 1;
}

static void throwing_handler(int status, const char* msg)
{
  std::fprintf(stderr, "axb error %s: %s\n", axb_status_string(status), msg);
  std::fflush(stdout);
  std::fflush(stderr);
  _exit(2);
}

int main(int argc, char** argv)
{
  const int N = argc > 1 ? std::atoi(argv[1]) : 1;
  const std::string dir = argc > 2 ? argv[2] : "/tmp";
  if(N < 1 || N > 16) return 64;
  ab::error_handler() = throwing_handler;
  const std::string id_file = dir + "/axb_nccl_id_" + std::to_string((long long)getpid());
  std::remove(id_file.c_str());
  // fork BEFORE anything touches CUDA: a CUDA context does not survive fork()
  std::vector<pid_t> kids;
  for(int r = 0; r < N; ++r)
  {
    const pid_t p = fork();
    if(p < 0) return 65;
    if(p == 0)
    {
      const int rc = run_rank(N, r, id_file);
      std::fflush(stdout);
      std::fflush(stderr);
      _exit(rc);
    }
    kids.push_back(p);
  }
  int fail = 0;
  for(pid_t p : kids)
  {
    int st = 0;
    waitpid(p, &st, 0);
    if(!WIFEXITED(st) || WEXITSTATUS(st) != 0) ++fail;
  }
  std::remove(id_file.c_str());
  if(fail == 0) std::printf("dcp_nccl_test: OK (%d ranks)\n", N);
  return fail == 0 ? 0 : 1;
}

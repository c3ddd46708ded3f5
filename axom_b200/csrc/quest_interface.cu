// quest_interface.cu -- the reference's legacy process-global signed-distance surface (host code only).
//
// Reference path replaced (paths relative to /root/reference/src/axom):
//   quest/interface/signed_distance.cpp:57-510   parameters, init(file) / init(mesh), setters, evaluate x3, finalize
//   quest/interface/c_fortran/wrapQUEST.cpp       the QUEST_signed_distance_* C symbols over them
//   quest/readers/STLReader.cpp:44-259            ASCII / binary STL -> triangle soup
//   quest/interface/internal/QuestHelpers.cpp:287-330, :486-514   read_stl_mesh, compute_mesh_bounds
//   quest/MeshTester.cpp:15-49, :218-333          weldTriMeshVertices
// All compute goes through the C ABI of api.cu (axb_sd_*): there is no CPU evaluation path here.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/axb200.h"
#include "../../include/axb200_quest.h"

namespace
{
constexpr int INIT_FAILED = -1;   // signed_distance.cpp:29
constexpr int INIT_SUCCESS = 0;   // :30

void default_handler(const char* msg)
{
  std::fprintf(stderr, "[ERROR] %s\n", msg);
  std::abort();  // slic abort-on-error is on by default (slic/core/Logger.cpp:37,51)
}
axb_quest_error_handler g_handler = default_handler;
void quest_error(const std::string& msg) { g_handler(msg.c_str()); }
void quest_warning(const std::string& msg) { std::fprintf(stderr, "[WARNING] %s\n", msg.c_str()); }

// parameters_t (signed_distance.cpp:57-86)
struct Parameters
{
  int dimension = 3;
  bool verbose = false;
  bool is_closed_surface = true;
  bool use_shared_memory = false;
  bool compute_sign = true;
  int allocator_id = -1;  // -1: device 0
  int exec_space = 0;     // SignedDistExec::CPU; every value is served by the GPU engine
} P;

axb_sd* s_query = nullptr;
double s_lo[3], s_hi[3];  // compute_mesh_bounds of the surface mesh

bool initialized() { return s_query != nullptr; }

// The whole file in one read; the format test and both parsers work on the buffer.
bool stl_slurp(const std::string& file, std::vector<char>& buf)
{
  FILE* f = std::fopen(file.c_str(), "rb");
  if(!f)
  {
    quest_warning("Cannot open the provided STL file [" + file + "]");
    return false;
  }
  std::fseek(f, 0, SEEK_END);
  const long size = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  buf.resize(size > 0 ? (size_t)size : 0);
  const size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
  std::fclose(f);
  buf.resize(got);
  return true;
}

// The reader's format rule (quest/readers/STLReader.cpp, isAsciiFormat): a file is binary exactly when its size is the
// 84-byte header plus 50 bytes per face of the count stored at offset 80; anything shorter than a header is ASCII.
bool stl_buffer_is_binary(const std::vector<char>& buf, std::int32_t& nfaces)
{
  constexpr size_t kHeader = 80 + sizeof(std::int32_t);
  nfaces = 0;
  if(buf.size() < kHeader) return false;
  std::memcpy(&nfaces, buf.data() + 80, sizeof(nfaces));
  // the reader compares in 32-bit arithmetic (an int32 file size against header + 50 * count)
  return static_cast<std::int32_t>(buf.size()) == static_cast<std::int32_t>(kHeader) + nfaces * 50;
}

// ASCII: every whitespace-delimited token "vertex" is followed by three numbers, one node (readAsciiSTL); the rest
// of the grammar (solid / facet / normal / outer loop ...) is skipped.  A pointer scan with strtod instead of
// formatted stream extraction.
void stl_parse_ascii(std::vector<char>& buf, std::vector<double>& nodes)
{
  buf.push_back('\0');  // strtod needs a terminator
  const char* p = buf.data();
  const char* const end = p + buf.size() - 1;
  auto is_space = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; };
  while(p < end)
  {
    while(p < end && is_space(*p)) ++p;
    const char* tok = p;
    while(p < end && !is_space(*p)) ++p;
    if(p - tok != 6 || std::memcmp(tok, "vertex", 6) != 0) continue;
    double v[3] = {0.0, 0.0, 0.0};
    bool ok = true;
    for(int k = 0; k < 3 && ok; ++k)
    {
      char* next = nullptr;
      v[k] = std::strtod(p, &next);
      ok = next != p && next <= end;
      if(ok) p = next;
    }
    nodes.insert(nodes.end(), v, v + 3);
    if(!ok) break;  // a malformed vertex ends the read (the stream reader stops at its first failed extraction)
  }
}

// Binary: 50-byte face records after the header -- float normal[3] (ignored), float vertex[3][3], uint16 attribute --
// vertices widened to double (readBinarySTL)
void stl_parse_binary(const std::vector<char>& buf, std::int32_t nfaces, std::vector<double>& nodes)
{
  if(nfaces <= 0) return;
  nodes.resize((size_t)nfaces * 9);
  const char* rec = buf.data() + 84;
  for(std::int32_t i = 0; i < nfaces; ++i, rec += 50)
  {
    float v[9];
    std::memcpy(v, rec + 12, sizeof(v));
    for(int j = 0; j < 9; ++j) nodes[(size_t)i * 9 + j] = static_cast<double>(v[j]);
  }
}

// STLReader::read + getMesh (:194-259): SoA coordinates, implicit connectivity 3i, 3i+1, 3i+2
int stl_read(const std::string& file, std::vector<double>& x, std::vector<double>& y, std::vector<double>& z, std::vector<int32_t>& conn)
{
  if(file.empty()) return -1;
  std::vector<char> buf;
  if(!stl_slurp(file, buf)) return -1;
  std::vector<double> nodes;
  std::int32_t nfaces = 0;
  if(stl_buffer_is_binary(buf, nfaces))
    stl_parse_binary(buf, nfaces, nodes);
  else
    stl_parse_ascii(buf, nodes);
  const size_t nn = nodes.size() / 3;
  const size_t nf = nn / 3;
  x.resize(nn);
  y.resize(nn);
  z.resize(nn);
  for(size_t i = 0; i < nn; ++i)
  {
    x[i] = nodes[3 * i];
    y[i] = nodes[3 * i + 1];
    z[i] = nodes[3 * i + 2];
  }
  conn.resize(nf * 3);
  for(size_t i = 0; i < nf * 3; ++i) conn[i] = (int32_t)i;
  return 0;
}

struct Cell3
{
  std::int64_t c[3];
  bool operator==(const Cell3& o) const { return c[0] == o.c[0] && c[1] == o.c[1] && c[2] == o.c[2]; }
};
struct Cell3Hash  // MeshTester.cpp:229-236 (only the bucket order depends on it; the result does not)
{
  size_t operator()(const Cell3& p) const
  {
    size_t seed = std::hash<std::int64_t> {}(p.c[0]);
    for(int i = 1; i < 3; ++i) seed ^= std::hash<std::int64_t> {}(p.c[i]) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    return seed;
  }
};

// weldTriMeshVertices (MeshTester.cpp:218-333)
void weld(std::vector<double>& x, std::vector<double>& y, std::vector<double>& z, std::vector<int32_t>& conn, double eps)
{
  // compute_bounds(mesh).expand(eps) (:15-42, BoundingBox.hpp:534-544)
  double lo[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308};
  double hi[3] = {-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
  for(size_t i = 0; i < x.size(); ++i)
  {
    const double p[3] = {x[i], y[i], z[i]};
    for(int d = 0; d < 3; ++d)
    {
      if(p[d] < lo[d]) lo[d] = p[d];
      if(p[d] > hi[d]) hi[d] = p[d];
    }
  }
  for(int d = 0; d < 3; ++d)
  {
    lo[d] -= eps;
    hi[d] += eps;
    if(lo[d] > hi[d]) std::swap(lo[d], hi[d]);  // checkAndFixBounds
  }
  // RectangularLattice (spin/RectangularLattice.hpp:144-157,202-220): cell = floor((p - origin) * (1/spacing)),
  // spacing snapped to 0 (and inverse 0) below PRIMAL_TINY
  double spacing = eps;
  if(std::fabs(spacing - 0.0) <= 1e-50) spacing = 0.0;
  const double inv = spacing != 0.0 ? 1.0 / spacing : 0.0;
  const double offsets[2] = {0.0, eps / 2.};
  for(int pass = 0; pass < 2; ++pass)
  {
    const double origin[3] = {lo[0] - offsets[pass], lo[1] - offsets[pass], lo[2] - offsets[pass]};
    const size_t nv = x.size();
    std::unordered_map<Cell3, std::int64_t, Cell3Hash> index(nv);
    std::vector<int32_t> remap(nv);
    std::vector<double> nx, ny, nz;
    int unique = 0;
    for(size_t i = 0; i < nv; ++i)
    {
      const double p[3] = {x[i], y[i], z[i]};
      Cell3 cell;
      for(int d = 0; d < 3; ++d) cell.c[d] = static_cast<std::int64_t>(std::floor((p[d] - origin[d]) * inv));
      auto res = index.insert(std::make_pair(cell, (std::int64_t)unique));
      if(res.second)
      {
        ++unique;
        nx.push_back(p[0]);
        ny.push_back(p[1]);
        nz.push_back(p[2]);
      }
      remap[i] = static_cast<int32_t>(res.first->second);
    }
    std::vector<int32_t> nconn;
    nconn.reserve(conn.size());
    for(size_t t = 0; t + 2 < conn.size(); t += 3)
    {
      const int32_t a = remap[conn[t]], b = remap[conn[t + 1]], c = remap[conn[t + 2]];
      if(a != b && b != c && c != a)  // areTriangleIndicesDistinct (:44-49)
      {
        nconn.push_back(a);
        nconn.push_back(b);
        nconn.push_back(c);
      }
    }
    x.swap(nx);
    y.swap(ny);
    z.swap(nz);
    conn.swap(nconn);
  }
}

int init_from_arrays(const double* x, const double* y, const double* z, int32_t nn, const int32_t* conn, int32_t nc, int memspace)
{
  if(initialized())
  {
    quest_error("signed distance query has already been initialized!");
    return INIT_FAILED;
  }
  const int device = P.allocator_id >= 0 ? P.allocator_id : 0;
  axb_sd* sd = nullptr;
  const int st = axb_sd_create(&sd, device, x, y, z, nn, conn, nullptr, nc, 3, memspace, P.is_closed_surface ? 1 : 0, P.compute_sign ? 1 : 0);
  if(st != AXB_OK)
  {
    quest_warning(std::string("signed distance initialisation failed: ") + axb_last_error());
    return INIT_FAILED;
  }
  axb_sd_get_mesh_bounds(sd, s_lo, s_hi);
  s_query = sd;
  return INIT_SUCCESS;
}

#define REQUIRE_NOT_INITIALIZED()                                                                                   \
  if(initialized())                                                                                                 \
  {                                                                                                                 \
    quest_error("signed distance query already initialized; setting option has no effect!");                       \
    return;                                                                                                         \
  }
}  // namespace

extern "C" {

void axb_quest_set_error_handler(axb_quest_error_handler h) { g_handler = h ? h : default_handler; }
void axb_host_free(void* p) { std::free(p); }

int axb_stl_read(const char* file, double** x, double** y, double** z, int32_t* num_nodes, int32_t** conn, int32_t* num_cells)
{
  if(!file || !x || !y || !z || !num_nodes || !conn || !num_cells) return -1;
  std::vector<double> vx, vy, vz;
  std::vector<int32_t> vc;
  if(stl_read(file, vx, vy, vz, vc) != 0) return -1;
  auto dup = [](const void* src, size_t bytes) {
    void* p = std::malloc(bytes ? bytes : 1);
    if(p && bytes) std::memcpy(p, src, bytes);
    return p;
  };
  *x = (double*)dup(vx.data(), vx.size() * sizeof(double));
  *y = (double*)dup(vy.data(), vy.size() * sizeof(double));
  *z = (double*)dup(vz.data(), vz.size() * sizeof(double));
  *conn = (int32_t*)dup(vc.data(), vc.size() * sizeof(int32_t));
  *num_nodes = (int32_t)vx.size();
  *num_cells = (int32_t)(vc.size() / 3);
  return 0;
}

int axb_weld_tri_mesh_vertices(double* x, double* y, double* z, int32_t* num_nodes, int32_t* conn, int32_t* num_cells, double eps)
{
  if(!x || !y || !z || !num_nodes || !conn || !num_cells || !(eps > 0.)) return -1;
  std::vector<double> vx(x, x + *num_nodes), vy(y, y + *num_nodes), vz(z, z + *num_nodes);
  std::vector<int32_t> vc(conn, conn + (size_t)*num_cells * 3);
  weld(vx, vy, vz, vc, eps);
  std::memcpy(x, vx.data(), vx.size() * sizeof(double));
  std::memcpy(y, vy.data(), vy.size() * sizeof(double));
  std::memcpy(z, vz.data(), vz.size() * sizeof(double));
  std::memcpy(conn, vc.data(), vc.size() * sizeof(int32_t));
  *num_nodes = (int32_t)vx.size();
  *num_cells = (int32_t)(vc.size() / 3);
  return 0;
}

// signed_distance_init(const std::string& file) (signed_distance.cpp:106-158)
int QUEST_signed_distance_init_serial(const char* file)
{
  if(P.dimension != 3)
  {
    quest_warning("the SignedDistance Query is currently only supported in 3D");
    return INIT_FAILED;
  }
  std::vector<double> x, y, z;
  std::vector<int32_t> conn;
  if(!file || stl_read(file, x, y, z, conn) != 0)
  {
    quest_warning(std::string("reading mesh from [") + (file ? file : "") + "] failed!");
    return INIT_FAILED;
  }
  return init_from_arrays(x.data(), y.data(), z.data(), (int32_t)x.size(), conn.data(), (int32_t)(conn.size() / 3), AXB_MEM_HOST);
}

int QUEST_signed_distance_init_serial_bufferify(char* file, int SHT_file_len)
{
  const std::string name(file ? file : "", file ? (size_t)(SHT_file_len > 0 ? SHT_file_len : 0) : 0);
  // Shroud trims the blank padding of a Fortran CHARACTER argument
  const size_t end = name.find_last_not_of(' ');
  return QUEST_signed_distance_init_serial(end == std::string::npos ? "" : name.substr(0, end + 1).c_str());
}

// signed_distance_init(const mint::Mesh* m) (:161-229)
int axb_quest_signed_distance_init_mesh(const double* x, const double* y, const double* z, int32_t nn, const int32_t* conn, int32_t nc,
                                        int memspace)
{
  return init_from_arrays(x, y, z, nn, conn, nc, memspace);
}

bool QUEST_signed_distance_initialized(void) { return initialized(); }

void QUEST_signed_distance_get_mesh_bounds(double* lo, double* hi)
{
  if(!initialized())
  {
    quest_error("signed distance query must be initialized prior to calling get_mesh_bounds()");
    return;
  }
  if(!lo || !hi)
  {
    quest_error("supplied buffer is null");
    return;
  }
  for(int d = 0; d < 3; ++d)
  {
    lo[d] = s_lo[d];
    hi[d] = s_hi[d];
  }
}

void QUEST_signed_distance_set_dimension(int dim)
{
  if(dim != 3)
  {
    quest_error("The signed distance query only support 3D");
    return;
  }
  REQUIRE_NOT_INITIALIZED();
  P.dimension = dim;
}
void QUEST_signed_distance_set_closed_surface(bool status)
{
  REQUIRE_NOT_INITIALIZED();
  P.is_closed_surface = status;
}
void QUEST_signed_distance_set_compute_signs(bool computeSign)
{
  REQUIRE_NOT_INITIALIZED();
  P.compute_sign = computeSign;
}
void QUEST_signed_distance_set_allocator(int allocatorID)
{
  REQUIRE_NOT_INITIALIZED();
  P.allocator_id = allocatorID;
}
void QUEST_signed_distance_set_verbose(bool status)
{
  REQUIRE_NOT_INITIALIZED();
  P.verbose = status;
}
void QUEST_signed_distance_use_shared_memory(bool status)
{
  REQUIRE_NOT_INITIALIZED();
  P.use_shared_memory = status;
  if(status) quest_warning("Enabling shared memory requires MPI-3. Option is ignored!");
}
void QUEST_signed_distance_set_execution_space(int execSpace)
{
  REQUIRE_NOT_INITIALIZED();
  if(execSpace < 0 || execSpace > 2)
  {
    quest_error("Unsupported execution space");
    return;
  }
  P.exec_space = execSpace;
}

// signed_distance_evaluate(x, y, z) (:342-372)
double QUEST_signed_distance_evaluate_0(double x, double y, double z)
{
  if(!initialized())
  {
    quest_error("signed distance query must be initialized prior to calling evaluate()!");
    return 0.0;
  }
  const double q[3] = {x, y, z};
  axb_array_desc d;
  std::memset(&d, 0, sizeof(d));
  for(int c = 0; c < 3; ++c) d.comp[c] = q + c;
  d.stride_bytes = 24;
  d.ncomp = 3;
  d.memspace = AXB_MEM_HOST;
  double phi = 0.0;
  if(axb_sd_compute_distances(s_query, &d, 1, &phi, nullptr, nullptr, AXB_MEM_HOST) != AXB_OK) quest_error(axb_last_error());
  return phi;
}

// signed_distance_evaluate(x, y, z, cp_x, ..., n_x, ...) (:375-428)
double QUEST_signed_distance_evaluate_1(double x, double y, double z, double* cp_x, double* cp_y, double* cp_z, double* n_x, double* n_y,
                                        double* n_z)
{
  if(!initialized())
  {
    quest_error("signed distance query must be initialized prior to calling evaluate()!");
    return 0.0;
  }
  const double q[3] = {x, y, z};
  axb_array_desc d;
  std::memset(&d, 0, sizeof(d));
  for(int c = 0; c < 3; ++c) d.comp[c] = q + c;
  d.stride_bytes = 24;
  d.ncomp = 3;
  d.memspace = AXB_MEM_HOST;
  double phi = 0.0, cp[3] = {0, 0, 0}, n[3] = {0, 0, 0};
  if(axb_sd_compute_distances(s_query, &d, 1, &phi, cp, n, AXB_MEM_HOST) != AXB_OK) quest_error(axb_last_error());
  if(cp_x) *cp_x = cp[0];
  if(cp_y) *cp_y = cp[1];
  if(cp_z) *cp_z = cp[2];
  if(n_x) *n_x = n[0];
  if(n_y) *n_y = n[1];
  if(n_z) *n_z = n[2];
  return phi;
}

// signed_distance_evaluate(const double* x, y, z, int npoints, double* phi) (:431-465): ZipIndexable SoA query
void axb_quest_signed_distance_evaluate_n(const double* x, const double* y, const double* z, int npoints, double* phi)
{
  if(!initialized())
  {
    quest_error("signed distance query must be initialized prior to calling evaluate()!");
    return;
  }
  if(!x)
  {
    quest_error("x-coords array is null");
    return;
  }
  if(!y)
  {
    quest_error("y-coords array is null");
    return;
  }
  if(!z)
  {
    quest_error("z-coords array is null");
    return;
  }
  if(!phi)
  {
    quest_error("output phi array is null");
    return;
  }
  axb_array_desc d;
  std::memset(&d, 0, sizeof(d));
  d.comp[0] = x;
  d.comp[1] = y;
  d.comp[2] = z;
  d.stride_bytes = 8;
  d.ncomp = 3;
  d.memspace = AXB_MEM_AUTO;
  if(axb_sd_compute_distances(s_query, &d, npoints, phi, nullptr, nullptr, AXB_MEM_AUTO) != AXB_OK) quest_error(axb_last_error());
}

// signed_distance_finalize (:468-507)
void QUEST_signed_distance_finalize(void)
{
  if(s_query)
  {
    axb_sd_destroy(s_query);
    s_query = nullptr;
  }
}

}  // extern "C"

// axom_b200/signed_distance.hpp -- the reference's process-global signed-distance interface
// (quest/interface/signed_distance.hpp:88-319) with the same function names and argument meaning, as inline
// forwards to the C symbols of axb200_quest.h.  With AXOM_B200_ALIAS_AXOM defined, `axom::quest::signed_distance_*`
// resolves here, so a host written against the reference only changes its include line and link line.
//
// Differences a maintainer should know (also in INTEGRATION.md): there is no MPI_Comm argument (one process per GPU;
// the surface is replicated), set_allocator takes the CUDA device ordinal, every SignedDistExec value runs on the
// GPU, and SLIC_ERROR conditions go to axb_quest_set_error_handler (default: print + abort, as slic does).
#ifndef AXOM_B200_SIGNED_DISTANCE_INTERFACE_HPP_
#define AXOM_B200_SIGNED_DISTANCE_INTERFACE_HPP_

#include <string>

#include "../axb200_quest.h"
#include "SignedDistance.hpp"

namespace axom_b200
{
namespace quest
{
enum class SignedDistExec  // signed_distance.hpp:88-93
{
  CPU = 0,
  OpenMP = 1,
  GPU = 2
};

inline int signed_distance_init(const std::string& file) { return QUEST_signed_distance_init_serial(file.c_str()); }
// signed_distance_init(const mint::Mesh*): the mesh extract of SignedDistance.hpp (triangles only, :171-180)
inline int signed_distance_init(const SurfaceMesh* m)
{
  if(!m || m->nodes_per_cell != 3 || m->cell_node_offsets) return -1;
  return axb_quest_signed_distance_init_mesh(m->x, m->y, m->z, m->num_nodes, m->cells_to_nodes, m->num_cells, AXB_MEM_AUTO);
}
inline bool signed_distance_initialized() { return QUEST_signed_distance_initialized(); }
inline void signed_distance_get_mesh_bounds(double* lo, double* hi) { QUEST_signed_distance_get_mesh_bounds(lo, hi); }
inline void signed_distance_set_dimension(int dim) { QUEST_signed_distance_set_dimension(dim); }
inline void signed_distance_set_closed_surface(bool status) { QUEST_signed_distance_set_closed_surface(status); }
inline void signed_distance_set_compute_signs(bool computeSign) { QUEST_signed_distance_set_compute_signs(computeSign); }
inline void signed_distance_set_allocator(int allocatorID) { QUEST_signed_distance_set_allocator(allocatorID); }
inline void signed_distance_set_verbose(bool status) { QUEST_signed_distance_set_verbose(status); }
inline void signed_distance_use_shared_memory(bool status) { QUEST_signed_distance_use_shared_memory(status); }
inline void signed_distance_set_execution_space(SignedDistExec e) { QUEST_signed_distance_set_execution_space(static_cast<int>(e)); }
inline double signed_distance_evaluate(double x, double y, double z = 0.0) { return QUEST_signed_distance_evaluate_0(x, y, z); }
inline double signed_distance_evaluate(double x, double y, double z, double& cp_x, double& cp_y, double& cp_z, double& n_x, double& n_y,
                                       double& n_z)
{
  return QUEST_signed_distance_evaluate_1(x, y, z, &cp_x, &cp_y, &cp_z, &n_x, &n_y, &n_z);
}
inline void signed_distance_evaluate(const double* x, const double* y, const double* z, int npoints, double* phi)
{
  axb_quest_signed_distance_evaluate_n(x, y, z, npoints, phi);
}
inline void signed_distance_finalize() { QUEST_signed_distance_finalize(); }

}  // namespace quest
}  // namespace axom_b200

#endif  // AXOM_B200_SIGNED_DISTANCE_INTERFACE_HPP_

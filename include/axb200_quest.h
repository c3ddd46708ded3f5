/* axb200_quest.h -- the reference's EXISTING C surface for the signed-distance path, served by the B200 engine.
 *
 * These are the symbols a C / Fortran / Python host of LLNL/axom already binds (Shroud wrappers,
 * quest/interface/c_fortran/wrapQUEST.h:83-127, over the process-global C++ functions of
 * quest/interface/signed_distance.hpp:117-319).  libaxb200.so exports them with the same names, argument meaning
 * and error behaviour, so such a host re-links against libaxb200.so instead of libaxom's quest and runs
 * on the GPU.  State is process-global and not thread-safe, exactly like the reference
 * (quest/interface/signed_distance.cpp:57-93).
 *
 * Error behaviour.  The reference reports misuse (evaluate before init, setters after init, null buffers) through
 * SLIC_ERROR, which logs and aborts while slic's abort-on-error is on -- the default (slic/core/Logger.cpp:37,51;
 * death tests in quest/tests/quest_signed_distance_interface.cpp:245-330).  Here the message goes to the handler set
 * with axb_quest_set_error_handler; the default handler prints it to stderr and calls abort().  A handler that
 * returns makes the call a no-op (evaluate returns 0.0), like slic with abort-on-error switched off.
 * init returns 0 on success and -1 on failure (signed_distance.cpp:29-30).
 */
#ifndef AXB200_QUEST_H_
#define AXB200_QUEST_H_

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- wrapQUEST.h:91-127 (serial build: AXOM_USE_MPI off) -------------------------------------------------- */
int QUEST_signed_distance_init_serial(const char* file);                         /* STL file (ASCII or binary)      */
int QUEST_signed_distance_init_serial_bufferify(char* file, int SHT_file_len);   /* Fortran: length-delimited name  */
bool QUEST_signed_distance_initialized(void);
void QUEST_signed_distance_get_mesh_bounds(double* lo, double* hi);
void QUEST_signed_distance_set_dimension(int dim);                               /* only 3 is supported             */
void QUEST_signed_distance_set_closed_surface(bool status);                      /* default true                    */
void QUEST_signed_distance_set_compute_signs(bool computeSign);                  /* default true                    */
void QUEST_signed_distance_set_allocator(int allocatorID);                       /* here: the CUDA device ordinal   */
void QUEST_signed_distance_set_verbose(bool status);
void QUEST_signed_distance_use_shared_memory(bool status);                       /* MPI-3 only: accepted, ignored   */
void QUEST_signed_distance_set_execution_space(int execSpace);                   /* SignedDistExec: 0 CPU 1 OpenMP 2 GPU;
                                                                                    every value runs on the B200     */
double QUEST_signed_distance_evaluate_0(double x, double y, double z);
double QUEST_signed_distance_evaluate_1(double x, double y, double z, double* cp_x, double* cp_y, double* cp_z, double* n_x,
                                        double* n_y, double* n_z);
void QUEST_signed_distance_finalize(void);

/* ---- the two C++ overloads Shroud does not wrap (signed_distance.hpp:117-150, :262-284) ------------------- */
/* signed_distance_init(const mint::Mesh*): a 3-D single-shape triangle mesh given by its arrays; the arrays are
 * copied to the device (the reference keeps the pointer).  memspace: axb_memspace of the mesh arrays. */
int axb_quest_signed_distance_init_mesh(const double* x, const double* y, const double* z, int32_t num_nodes,
                                        const int32_t* triangles_to_nodes, int32_t num_cells, int memspace);
/* signed_distance_evaluate(const double* x, y, z, int npoints, double* phi): the batched query; x, y, z, phi may be
 * host or device arrays (all in the same space). */
void axb_quest_signed_distance_evaluate_n(const double* x, const double* y, const double* z, int npoints, double* phi);

/* ---- input side of the path: STL ingestion and vertex welding --------------------------------------------- */
/* quest::STLReader::read + getMesh (quest/readers/STLReader.cpp:44-259): ASCII or binary STL -> triangle soup,
 * node 3i+k is vertex k of triangle i.  Arrays are malloc'ed; release with axb_host_free.  Returns 0 / -1. */
int axb_stl_read(const char* file, double** x, double** y, double** z, int32_t* num_nodes, int32_t** triangles_to_nodes,
                 int32_t* num_cells);
/* quest::weldTriMeshVertices(&mesh, eps) (quest/MeshTester.cpp:218-333): vertices that fall into the same cell of
 * an eps-lattice are merged (two passes, the second on a lattice shifted by eps/2), triangles that lose a vertex are
 * dropped.  In place: the arrays are compacted and the counts updated; numbering is the reference's (first
 * appearance).  Returns 0 / -1. */
int axb_weld_tri_mesh_vertices(double* x, double* y, double* z, int32_t* num_nodes, int32_t* triangles_to_nodes, int32_t* num_cells,
                               double eps);
void axb_host_free(void* p);

typedef void (*axb_quest_error_handler)(const char* message);
void axb_quest_set_error_handler(axb_quest_error_handler handler); /* NULL restores print + abort */

#ifdef __cplusplus
}
#endif
#endif /* AXB200_QUEST_H_ */

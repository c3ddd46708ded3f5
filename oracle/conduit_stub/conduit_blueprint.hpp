// conduit_blueprint.hpp -- MOCK (see conduit_node.hpp): the few blueprint predicates the reference's marching-cubes path asks.
#pragma once
#include "conduit_node.hpp"

namespace conduit
{
namespace blueprint
{
inline bool verify(const std::string&, const Node&, Node&) { return true; }
inline bool is_contiguous(const Node&) { return true; }
namespace mcarray
{
inline bool is_interleaved(const Node&) { return false; }  // the driver always passes one array per component
}
namespace mesh
{
inline bool is_multi_domain(const Node& n) { return !n.has_child("coordsets"); }
inline index_t number_of_domains(const Node& n) { return is_multi_domain(n) ? n.number_of_children() : 1; }
namespace coordset
{
inline index_t dims(const Node& cs) { return cs.fetch_existing("values").number_of_children(); }
}
namespace topology
{
// structured topology: the number of entries among elements/dims/{i,j,k}
inline index_t dims(const Node& topo)
{
  const Node& d = topo.fetch_existing("elements/dims");
  return (index_t)d.has_child("i") + (index_t)d.has_child("j") + (index_t)d.has_child("k");
}
}  // namespace topology
}  // namespace mesh
}  // namespace blueprint
}  // namespace conduit

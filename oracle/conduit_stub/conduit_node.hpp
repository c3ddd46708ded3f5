// conduit_node.hpp -- a MOCK of the tiny part of Conduit's Node API that the reference's quest::MarchingCubes /
// quest::MeshViewUtil touch.  TEST INFRASTRUCTURE ONLY (oracle/build_ref.py): Conduit is an external library that
// is not in this image; with this mock the UNMODIFIED reference sources quest/MarchingCubes.cpp,
// quest/detail/MarchingCubesSingleDomain.cpp, quest/detail/MarchingCubesImpl.hpp and quest/MeshViewUtil.hpp compile
// where they lie, so the marching-cubes oracle can be pinned to the real reference.  Nothing here computes anything:
// a Node is a named tree whose leaves are a string or an EXTERNAL typed array (pointer + element count).
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace conduit
{
using index_t = long;

class DataType
{
public:
  enum Kind { EMPTY, INT32, INT64, FLOAT64, STRING };
  DataType() = default;
  DataType(Kind k, index_t n) : m_kind(k), m_n(n) { }
  static DataType int32(index_t n = 1) { return DataType(INT32, n); }
  static DataType int64(index_t n = 1) { return DataType(INT64, n); }
  static DataType float64(index_t n = 1) { return DataType(FLOAT64, n); }
  index_t number_of_elements() const { return m_n; }
  void set_number_of_elements(index_t n) { m_n = n; }
  bool is_int32() const { return m_kind == INT32; }
  bool is_int64() const { return m_kind == INT64; }
  bool is_float64() const { return m_kind == FLOAT64; }
  Kind kind() const { return m_kind; }
  size_t element_bytes() const { return m_kind == INT32 ? 4 : 8; }

private:
  Kind m_kind = EMPTY;
  index_t m_n = 0;
};

class Node
{
public:
  // a value convertible to any integer type, as Node::to_value() returns in Conduit
  struct Value
  {
    long long v;
    template <typename T>
    operator T() const
    {
      return static_cast<T>(v);
    }
  };

  Node() = default;
  Node(const Node&) = delete;
  Node& operator=(const Node&) = delete;

  const std::string& name() const { return m_name; }

  // ---- tree -------------------------------------------------------------------------
  Node& fetch(const std::string& path)
  {
    Node* n = this;
    size_t b = 0;
    while(b <= path.size())
    {
      size_t e = path.find('/', b);
      if(e == std::string::npos) e = path.size();
      const std::string key = path.substr(b, e - b);
      if(!key.empty()) n = &n->child_or_create(key);
      b = e + 1;
    }
    return *n;
  }
  Node& operator[](const std::string& path) { return fetch(path); }
  const Node& operator[](const std::string& path) const { return fetch_existing(path); }
  Node& operator[](int i) { return *m_children.at((size_t)i); }
  const Node& operator[](int i) const { return *m_children.at((size_t)i); }
  Node& child(int i) { return *m_children.at((size_t)i); }
  const Node& child(int i) const { return *m_children.at((size_t)i); }
  index_t number_of_children() const { return (index_t)m_children.size(); }

  const Node* find(const std::string& path) const
  {
    const Node* n = this;
    size_t b = 0;
    while(b <= path.size() && n)
    {
      size_t e = path.find('/', b);
      if(e == std::string::npos) e = path.size();
      const std::string key = path.substr(b, e - b);
      if(!key.empty())
      {
        auto it = n->m_index.find(key);
        n = it == n->m_index.end() ? nullptr : n->m_children[it->second].get();
      }
      b = e + 1;
    }
    return n;
  }
  bool has_path(const std::string& path) const { return find(path) != nullptr; }
  bool has_child(const std::string& name) const { return m_index.count(name) != 0; }
  const Node& fetch_existing(const std::string& path) const
  {
    const Node* n = find(path);
    if(!n) throw std::runtime_error("conduit mock: no such path " + path);
    return *n;
  }
  Node& fetch_existing(const std::string& path) { return const_cast<Node&>(static_cast<const Node*>(this)->fetch_existing(path)); }

  // ---- leaves -----------------------------------------------------------------------
  void set(const std::string& s)
  {
    m_string = s;
    m_dtype = DataType(DataType::STRING, (index_t)s.size());
  }
  Node& operator=(const std::string& s)
  {
    set(s);
    return *this;
  }
  Node& operator=(const char* s)
  {
    set(std::string(s));
    return *this;
  }
  void set_external(const DataType& t, void* p)
  {
    m_dtype = t;
    m_data = p;
  }
  // owning variants (MeshViewUtil::createField; unused by MarchingCubes but they must compile)
  void set(const DataType& t)
  {
    m_owned.assign((size_t)t.number_of_elements() * t.element_bytes(), 0);
    m_dtype = t;
    m_data = m_owned.data();
  }
  void set(const DataType& t, const void* p)
  {
    set(t);
    std::memcpy(m_data, p, m_owned.size());
  }
  void set_int32(int32_t v)
  {
    set(DataType::int32(1));
    *static_cast<int32_t*>(m_data) = v;
  }

  const DataType& dtype() const { return m_dtype; }
  std::string as_string() const { return m_string; }
  double* as_double_ptr() const { return static_cast<double*>(m_data); }
  int32_t* as_int32_ptr() const { return static_cast<int32_t*>(m_data); }
  int64_t* as_int64_ptr() const { return static_cast<int64_t*>(m_data); }
  void* data_ptr() { return m_data; }
  const void* data_ptr() const { return m_data; }
  long long to_ll() const
  {
    if(m_dtype.is_int32()) return *static_cast<const int32_t*>(m_data);
    if(m_dtype.is_int64()) return *static_cast<const int64_t*>(m_data);
    throw std::runtime_error("conduit mock: to_value on a non-integer leaf");
  }
  Value to_value() const { return Value {to_ll()}; }
  int32_t to_int32() const { return (int32_t)to_ll(); }
  void print() const { }

private:
  Node& child_or_create(const std::string& key)
  {
    auto it = m_index.find(key);
    if(it != m_index.end()) return *m_children[it->second];
    m_index[key] = m_children.size();
    m_children.emplace_back(new Node());
    m_children.back()->m_name = key;
    return *m_children.back();
  }
  std::string m_name;
  std::vector<std::unique_ptr<Node>> m_children;  // insertion order = child index, as in Conduit
  std::map<std::string, size_t> m_index;
  DataType m_dtype;
  void* m_data = nullptr;
  std::string m_string;
  std::vector<char> m_owned;
};
}  // namespace conduit

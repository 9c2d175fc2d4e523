// Minimal stand-in for <boost/graph/adjacency_list.hpp> (see oracle/shim/README.md).
// Only adjacency_list<setS, setS, undirectedS, VP, EP> and the free functions the reference's
// ift.hpp / globals.hpp call.  Written from the Boost.Graph API documentation; test infrastructure.
#ifndef IFTWT_SHIM_BOOST_ADJACENCY_LIST_HPP
#define IFTWT_SHIM_BOOST_ADJACENCY_LIST_HPP
#include <algorithm>
#include <cstddef>
#include <functional>
#include <deque>
#include <iterator>
#include <list>
#include <memory>
#include <set>
#include <utility>
#include <vector>

namespace boost {
struct setS {};
struct vecS {};
struct listS {};
struct undirectedS {};
struct directedS {};
struct no_property {};

template <class It>
struct iterator_range {
  It b, e;
  It begin() const { return b; }
  It end() const { return e; }
};
template <class It>
iterator_range<It> make_iterator_range(const std::pair<It, It>& p) {
  return iterator_range<It>{p.first, p.second};
}

template <class OutEdgeS, class VertexS, class DirS, class VP = no_property, class EP = no_property>
class adjacency_list {
 public:
  typedef void* vertex_descriptor;
  typedef VP vertex_property_type;
  typedef EP edge_property_type;
  struct edge_record {
    vertex_descriptor s, t;
    EP prop;
  };
  struct edge_descriptor {
    vertex_descriptor s, t;
    edge_record* rec;
    bool operator==(const edge_descriptor& o) const { return rec == o.rec; }
  };
  struct out_edge {
    vertex_descriptor target;
    edge_record* rec;
    bool operator<(const out_edge& o) const { return std::less<void*>()(target, o.target); }  // setS: by target address
  };
  struct node {
    VP prop;
    std::set<out_edge> out;
    std::size_t index;  // creation order inside this graph (shim extension, used by the driver only)
  };

  class vertex_iterator {
   public:
    typedef std::forward_iterator_tag iterator_category;
    typedef vertex_descriptor value_type;
    typedef std::ptrdiff_t difference_type;
    typedef const vertex_descriptor* pointer;
    typedef vertex_descriptor reference;
    vertex_iterator() {}
    explicit vertex_iterator(std::set<void*>::const_iterator i) : it(i) {}
    vertex_descriptor operator*() const { return *it; }
    vertex_iterator& operator++() { ++it; return *this; }
    vertex_iterator operator++(int) { vertex_iterator t = *this; ++it; return t; }
    bool operator==(const vertex_iterator& o) const { return it == o.it; }
    bool operator!=(const vertex_iterator& o) const { return it != o.it; }
   private:
    std::set<void*>::const_iterator it;
  };
  class adjacency_iterator {
   public:
    typedef std::forward_iterator_tag iterator_category;
    typedef vertex_descriptor value_type;
    typedef std::ptrdiff_t difference_type;
    typedef const vertex_descriptor* pointer;
    typedef const vertex_descriptor& reference;
    adjacency_iterator() {}
    explicit adjacency_iterator(typename std::set<out_edge>::const_iterator i) : it(i) {}
    const vertex_descriptor& operator*() const { return it->target; }
    adjacency_iterator& operator++() { ++it; return *this; }
    adjacency_iterator operator++(int) { adjacency_iterator t = *this; ++it; return t; }
    bool operator==(const adjacency_iterator& o) const { return it == o.it; }
    bool operator!=(const adjacency_iterator& o) const { return it != o.it; }
   private:
    typename std::set<out_edge>::const_iterator it;
  };
  class edge_iterator {
   public:
    typedef std::forward_iterator_tag iterator_category;
    typedef edge_descriptor value_type;
    typedef std::ptrdiff_t difference_type;
    typedef const edge_descriptor* pointer;
    typedef edge_descriptor reference;
    edge_iterator() {}
    explicit edge_iterator(typename std::list<edge_record>::iterator i) : it(i) {}
    edge_descriptor operator*() const { return edge_descriptor{it->s, it->t, &*it}; }
    edge_iterator& operator++() { ++it; return *this; }
    edge_iterator operator++(int) { edge_iterator t = *this; ++it; return t; }
    bool operator==(const edge_iterator& o) const { return it == o.it; }
    bool operator!=(const edge_iterator& o) const { return it != o.it; }
   private:
    typename std::list<edge_record>::iterator it;
  };

  adjacency_list() {}
  adjacency_list(const adjacency_list& o) { copy_from(o); }
  adjacency_list& operator=(const adjacency_list& o) {
    if (this != &o) {
      clear();
      copy_from(o);
    }
    return *this;
  }
  void clear() {
    vset_.clear();
    edges_.clear();
    arena_.clear();
  }

  VP& operator[](vertex_descriptor v) { return static_cast<node*>(v)->prop; }
  const VP& operator[](vertex_descriptor v) const { return static_cast<const node*>(v)->prop; }
  EP& operator[](edge_descriptor e) { return e.rec->prop; }

  // --- used by the free functions below ---
  vertex_descriptor add_vertex_(const VP& p) {
    if (arena_.empty() || arena_.back().size() == arena_.back().capacity()) {
      // Chunks are reserved up front, so nodes never move and inside a chunk addresses ascend with
      // creation order.  The first chunk holds 4 M nodes (untouched pages are never committed), so
      // for every graph the tests build address order IS creation order.
      arena_.emplace_back();
      arena_.back().reserve((std::size_t)1 << 22);
    }
    arena_.back().push_back(node{p, {}, count_++});
    node* n = &arena_.back().back();
    vset_.insert(n);
    return n;
  }
  std::pair<edge_descriptor, bool> add_edge_(vertex_descriptor u, vertex_descriptor v, const EP& p) {
    node* nu = static_cast<node*>(u);
    node* nv = static_cast<node*>(v);
    auto f = nu->out.find(out_edge{v, nullptr});
    if (f != nu->out.end()) return std::make_pair(edge_descriptor{u, v, f->rec}, false);  // setS: no parallel edges
    edges_.push_back(edge_record{u, v, p});
    edge_record* r = &edges_.back();
    nu->out.insert(out_edge{v, r});
    if (u != v) nv->out.insert(out_edge{u, r});
    return std::make_pair(edge_descriptor{u, v, r}, true);
  }
  const std::set<void*>& vset() const { return vset_; }
  std::list<edge_record>& edge_list() { return edges_; }
  const std::list<edge_record>& edge_list() const { return edges_; }
  static std::size_t index_of(vertex_descriptor v) { return static_cast<node*>(v)->index; }

 private:
  void copy_from(const adjacency_list& o) {
    count_ = 0;
    // vertices in the source's iteration order, edges in its edge-list order
    std::vector<std::pair<void*, void*>> map;
    map.reserve(o.vset_.size());
    for (void* v : o.vset_) map.emplace_back(v, add_vertex_(static_cast<node*>(v)->prop));
    auto find = [&](void* v) {
      auto it = std::lower_bound(map.begin(), map.end(), std::make_pair(v, (void*)nullptr),
                                 [](const std::pair<void*, void*>& a, const std::pair<void*, void*>& b) {
                                   return std::less<void*>()(a.first, b.first);
                                 });
      return it->second;
    };
    for (const edge_record& e : o.edges_) add_edge_(find(e.s), find(e.t), e.prop);
  }
  // one graph's nodes live in a few large chunks that never reallocate
  std::deque<std::vector<node>> arena_;
  std::set<void*> vset_;
  std::list<edge_record> edges_;
  std::size_t count_ = 0;
};

template <class G>
struct graph_traits {
  typedef typename G::vertex_descriptor vertex_descriptor;
  typedef typename G::edge_descriptor edge_descriptor;
  typedef typename G::vertex_iterator vertex_iterator;
  typedef typename G::edge_iterator edge_iterator;
  typedef typename G::adjacency_iterator adjacency_iterator;
};

#define IFTWT_SHIM_AL adjacency_list<OE, VL, D, VP, EP>
#define IFTWT_SHIM_TP template <class OE, class VL, class D, class VP, class EP>

IFTWT_SHIM_TP typename IFTWT_SHIM_AL::vertex_descriptor add_vertex(const VP& p, IFTWT_SHIM_AL& g) { return g.add_vertex_(p); }
IFTWT_SHIM_TP typename IFTWT_SHIM_AL::vertex_descriptor add_vertex(IFTWT_SHIM_AL& g) { return g.add_vertex_(VP()); }
IFTWT_SHIM_TP std::pair<typename IFTWT_SHIM_AL::edge_descriptor, bool> add_edge(void* u, void* v, IFTWT_SHIM_AL& g) {
  return g.add_edge_(u, v, EP());
}
IFTWT_SHIM_TP std::pair<typename IFTWT_SHIM_AL::edge_descriptor, bool> add_edge(void* u, void* v, const EP& p,
                                                                             IFTWT_SHIM_AL& g) {
  return g.add_edge_(u, v, p);
}
IFTWT_SHIM_TP std::pair<typename IFTWT_SHIM_AL::vertex_iterator, typename IFTWT_SHIM_AL::vertex_iterator> vertices(
    const IFTWT_SHIM_AL& g) {
  typedef typename IFTWT_SHIM_AL::vertex_iterator It;
  return std::make_pair(It(g.vset().begin()), It(g.vset().end()));
}
IFTWT_SHIM_TP std::pair<typename IFTWT_SHIM_AL::edge_iterator, typename IFTWT_SHIM_AL::edge_iterator> edges(
    IFTWT_SHIM_AL& g) {
  typedef typename IFTWT_SHIM_AL::edge_iterator It;
  return std::make_pair(It(g.edge_list().begin()), It(g.edge_list().end()));
}
IFTWT_SHIM_TP std::pair<typename IFTWT_SHIM_AL::adjacency_iterator, typename IFTWT_SHIM_AL::adjacency_iterator>
adjacent_vertices(void* v, const IFTWT_SHIM_AL&) {
  typedef typename IFTWT_SHIM_AL::adjacency_iterator It;
  typedef typename IFTWT_SHIM_AL::node node;
  const node* n = static_cast<const node*>(v);
  return std::make_pair(It(n->out.begin()), It(n->out.end()));
}
IFTWT_SHIM_TP void* source(const typename IFTWT_SHIM_AL::edge_descriptor& e, const IFTWT_SHIM_AL&) { return e.s; }
IFTWT_SHIM_TP void* target(const typename IFTWT_SHIM_AL::edge_descriptor& e, const IFTWT_SHIM_AL&) { return e.t; }
IFTWT_SHIM_TP std::size_t num_vertices(const IFTWT_SHIM_AL& g) { return g.vset().size(); }
IFTWT_SHIM_TP std::size_t num_edges(const IFTWT_SHIM_AL& g) { return g.edge_list().size(); }

#undef IFTWT_SHIM_AL
#undef IFTWT_SHIM_TP
}  // namespace boost

// <boost/graph/copy.hpp>, as called at graph.hpp:315-317: vertices in the source's iteration order,
// edges in its edge-list order; the vertex-index map only serves Boost's internal bookkeeping
namespace boost {
template <class M>
struct shim_assoc_map {
  M* m;
};
template <class M>
shim_assoc_map<M> make_assoc_property_map(M& m) { return shim_assoc_map<M>{&m}; }
template <class P>
struct shim_named_param {
  P p;
};
template <class P>
shim_named_param<P> vertex_index_map(const P& p) { return shim_named_param<P>{p}; }
template <class G, class P>
void copy_graph(const G& src, G& dst, const shim_named_param<P>&) { dst = src; }
template <class G>
void copy_graph(const G& src, G& dst) { dst = src; }
}  // namespace boost

// <boost/graph/iteration_macros.hpp>
#define BGL_FORALL_VERTICES(VNAME, GNAME, GraphType) \
  for (GraphType::vertex_descriptor VNAME : boost::make_iterator_range(boost::vertices(GNAME)))
#define BGL_FORALL_EDGES(ENAME, GNAME, GraphType) \
  for (GraphType::edge_descriptor ENAME : boost::make_iterator_range(boost::edges(GNAME)))

#endif

// Minimal stand-in for the oneTBB API surface the reference's placement translation units use
// (src/WEPP/{initial_filter,arena,util,dataset}.cpp and the headers they include).
// TEST INFRASTRUCTURE: lets those files compile from where they lie under /root/reference
// without oneTBB (absent from this image).  parallel_for is a std::thread blocked-range pool
// honouring global_control(max_allowed_parallelism), so the shim-compiled reference still runs
// its own TBB decomposition over reads on all host cores.  Written from the documented oneTBB
// interface; no reference or oneTBB source is copied.
#pragma once
// oneTBB's own headers transitively provide these; the reference relies on that (INT_MAX, std::log,
// unqualified abs on doubles — src/WEPP/arena.hpp:21 — needs the <stdlib.h>/<math.h> overloads).
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <climits>
#include <cmath>
#include <cstring>
#include <iostream>
#include <map>
#include <queue>
#include <set>
#include <algorithm>
#include <atomic>
#include <cstddef>
#include <functional>
#include <iterator>
#include <list>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

namespace tbb {

class global_control {
  public:
    enum parameter { max_allowed_parallelism, thread_stack_size };
    global_control(parameter p, size_t value) : p_(p) {
        if (p == max_allowed_parallelism) {
            prev_ = limit();
            limit() = value ? value : 1;
        }
    }
    global_control(const global_control& o) : p_(o.p_), prev_(limit()) {}
    ~global_control() {}
    static size_t active_value(parameter p) {
        if (p == max_allowed_parallelism) return limit();
        return 0;
    }
  private:
    static size_t& limit() {
        static size_t v = std::max(1u, std::thread::hardware_concurrency());
        return v;
    }
    parameter p_;
    size_t prev_ = 0;
};

template <typename T>
class blocked_range {
  public:
    using const_iterator = T;
    blocked_range(T b, T e, size_t grain = 1) : b_(b), e_(e), grain_(grain ? grain : 1) {}
    T begin() const { return b_; }
    T end() const { return e_; }
    size_t size() const { return (size_t)(e_ - b_); }
    size_t grainsize() const { return grain_; }
    bool empty() const { return !(b_ < e_); }
  private:
    T b_, e_;
    size_t grain_;
};

struct auto_partitioner {};
struct simple_partitioner {};
struct static_partitioner {};
struct affinity_partitioner {};

namespace detail {
template <typename T, typename Body>
void run_chunks(const blocked_range<T>& r, const Body& body) {
    const size_t n = r.size();
    if (n == 0) return;
    size_t threads = global_control::active_value(global_control::max_allowed_parallelism);
    // chunking: like TBB's auto partitioner, split down towards the grain size but not below
    size_t chunk = std::max<size_t>(r.grainsize(), 1);
    if (r.grainsize() <= 1) chunk = std::max<size_t>(1, n / (threads * 8));
    const size_t n_chunks = (n + chunk - 1) / chunk;
    threads = std::min(threads, n_chunks);
    if (threads <= 1) {
        for (size_t c = 0; c < n_chunks; ++c) {
            T b = r.begin() + (T)(c * chunk);
            T e = r.begin() + (T)std::min(n, (c + 1) * chunk);
            body(blocked_range<T>(b, e, r.grainsize()));
        }
        return;
    }
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (;;) {
            size_t c = next.fetch_add(1);
            if (c >= n_chunks) break;
            T b = r.begin() + (T)(c * chunk);
            T e = r.begin() + (T)std::min(n, (c + 1) * chunk);
            body(blocked_range<T>(b, e, r.grainsize()));
        }
    };
    std::vector<std::thread> pool;
    for (size_t t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
}
}  // namespace detail

template <typename T, typename Body>
void parallel_for(const blocked_range<T>& r, const Body& body) { detail::run_chunks(r, body); }
template <typename T, typename Body, typename Part>
void parallel_for(const blocked_range<T>& r, const Body& body, Part&&) { detail::run_chunks(r, body); }

template <typename It>
void parallel_sort(It b, It e) { std::sort(b, e); }
template <typename It, typename Cmp>
void parallel_sort(It b, It e, Cmp c) { std::sort(b, e, c); }

class queuing_mutex {
  public:
    class scoped_lock {
      public:
        scoped_lock() = default;
        explicit scoped_lock(queuing_mutex& m) : m_(&m) { m_->m_.lock(); }
        ~scoped_lock() { release(); }
        void acquire(queuing_mutex& m) { m_ = &m; m_->m_.lock(); }
        void release() { if (m_) { m_->m_.unlock(); m_ = nullptr; } }
        scoped_lock(const scoped_lock&) = delete;
        scoped_lock& operator=(const scoped_lock&) = delete;
      private:
        queuing_mutex* m_ = nullptr;
    };
    queuing_mutex() = default;
    queuing_mutex(const queuing_mutex&) {}
  private:
    std::mutex m_;
};
using spin_mutex = queuing_mutex;
using mutex = queuing_mutex;

class queuing_rw_mutex {
  public:
    class scoped_lock {
      public:
        scoped_lock() = default;
        scoped_lock(queuing_rw_mutex& m, bool write = true) { acquire(m, write); }
        ~scoped_lock() { release(); }
        void acquire(queuing_rw_mutex& m, bool write = true) {
            m_ = &m; w_ = write;
            if (w_) m_->m_.lock(); else m_->m_.lock_shared();
        }
        void release() {
            if (!m_) return;
            if (w_) m_->m_.unlock(); else m_->m_.unlock_shared();
            m_ = nullptr;
        }
        bool upgrade_to_writer() { if (!w_) { m_->m_.unlock_shared(); m_->m_.lock(); w_ = true; } return false; }
        bool downgrade_to_reader() { if (w_) { m_->m_.unlock(); m_->m_.lock_shared(); w_ = false; } return false; }
      private:
        queuing_rw_mutex* m_ = nullptr;
        bool w_ = true;
    };
  private:
    std::shared_mutex m_;
};

// concurrent_unordered_{map,set}: std containers whose inserting members take a lock (element references of
// the node-based std containers stay valid across rehashing, as TBB's do), so the reference's concurrent
// emplace / operator[] from inside parallel_for (src/mutation_annotated_tree.cpp:598-610) is safe.
template <typename K, typename V, typename H = std::hash<K>, typename E = std::equal_to<K>>
class concurrent_unordered_map : public std::unordered_map<K, V, H, E> {
  public:
    using base = std::unordered_map<K, V, H, E>;
    concurrent_unordered_map() {}
    concurrent_unordered_map(const concurrent_unordered_map& o) : base(o) {}
    concurrent_unordered_map& operator=(const concurrent_unordered_map& o) { base::operator=(o); return *this; }
    template <typename... A> auto emplace(A&&... a) { std::lock_guard<std::mutex> g(m_); return base::emplace(std::forward<A>(a)...); }
    template <typename A> auto insert(A&& a) { std::lock_guard<std::mutex> g(m_); return base::insert(std::forward<A>(a)); }
    V& operator[](const K& k) { std::lock_guard<std::mutex> g(m_); return base::operator[](k); }
    void unsafe_erase(const K& k) { this->erase(k); }
  private:
    std::mutex m_;
};
template <typename K, typename H = std::hash<K>, typename E = std::equal_to<K>>
class concurrent_unordered_set : public std::unordered_set<K, H, E> {
  public:
    using base = std::unordered_set<K, H, E>;
    concurrent_unordered_set() {}
    concurrent_unordered_set(const concurrent_unordered_set& o) : base(o) {}
    concurrent_unordered_set& operator=(const concurrent_unordered_set& o) { base::operator=(o); return *this; }
    template <typename... A> auto emplace(A&&... a) { std::lock_guard<std::mutex> g(m_); return base::emplace(std::forward<A>(a)...); }
    template <typename A> auto insert(A&& a) { std::lock_guard<std::mutex> g(m_); return base::insert(std::forward<A>(a)); }
  private:
    std::mutex m_;
};

template <typename K, typename V>
class concurrent_hash_map {
    struct Slot {
        std::pair<const K, V> kv;
        std::mutex m;
        template <typename KK, typename VV>
        Slot(KK&& k, VV&& v) : kv(std::forward<KK>(k), std::forward<VV>(v)) {}
    };
    using Map = std::unordered_map<K, std::unique_ptr<Slot>>;
  public:
    using value_type = std::pair<const K, V>;
    class const_accessor {
      public:
        ~const_accessor() { release(); }
        void release() { if (s_) { s_->m.unlock(); s_ = nullptr; } }
        bool empty() const { return s_ == nullptr; }
        const value_type& operator*() const { return s_->kv; }
        const value_type* operator->() const { return &s_->kv; }
      protected:
        friend class concurrent_hash_map;
        Slot* s_ = nullptr;
    };
    class accessor : public const_accessor {
      public:
        value_type& operator*() const { return this->s_->kv; }
        value_type* operator->() const { return &this->s_->kv; }
    };
    class iterator {
      public:
        using iterator_category = std::forward_iterator_tag;
        using value_type = std::pair<const K, V>;
        using difference_type = std::ptrdiff_t;
        using pointer = value_type*;
        using reference = value_type&;
        iterator() = default;
        explicit iterator(typename Map::iterator it) : it_(it) {}
        iterator operator++(int) { iterator t = *this; ++it_; return t; }
        value_type& operator*() const { return it_->second->kv; }
        value_type* operator->() const { return &it_->second->kv; }
        iterator& operator++() { ++it_; return *this; }
        bool operator!=(const iterator& o) const { return it_ != o.it_; }
        bool operator==(const iterator& o) const { return it_ == o.it_; }
      private:
        typename Map::iterator it_;
    };
    iterator begin() { return iterator(map_.begin()); }
    iterator end() { return iterator(map_.end()); }
    size_t size() const { return map_.size(); }
    bool empty() const { return map_.empty(); }

    bool insert(const_accessor& a, const K& k) { return insert_impl(a, k, V()); }
    bool insert(const_accessor& a, const value_type& kv) { return insert_impl(a, kv.first, kv.second); }
    bool insert(const_accessor& a, const std::pair<K, V>& kv) { return insert_impl(a, kv.first, kv.second); }
    bool insert(const value_type& kv) { const_accessor a; return insert_impl(a, kv.first, kv.second); }
    bool find(const_accessor& a, const K& k) {
        a.release();
        Slot* s = nullptr;
        {
            std::lock_guard<std::mutex> g(m_);
            auto it = map_.find(k);
            if (it == map_.end()) return false;
            s = it->second.get();
        }
        s->m.lock();
        a.s_ = s;
        return true;
    }
    size_t count(const K& k) { std::lock_guard<std::mutex> g(m_); return map_.count(k); }
    bool erase(const K& k) { std::lock_guard<std::mutex> g(m_); return map_.erase(k) > 0; }
    void clear() { std::lock_guard<std::mutex> g(m_); map_.clear(); }
  private:
    template <typename VV>
    bool insert_impl(const_accessor& a, const K& k, VV&& v) {
        a.release();
        Slot* s = nullptr;
        bool created = false;
        {
            std::lock_guard<std::mutex> g(m_);
            auto it = map_.find(k);
            if (it == map_.end()) {
                it = map_.emplace(k, std::make_unique<Slot>(k, std::forward<VV>(v))).first;
                created = true;
            }
            s = it->second.get();
        }
        s->m.lock();
        a.s_ = s;
        return created;
    }
    Map map_;
    std::mutex m_;
};

template <typename T>
using scalable_allocator = std::allocator<T>;

class task_group {
  public:
    template <typename F> void run(F&& f) { f(); }
    void wait() {}
};

// flow graph: only has to compile (MAT::read_vcf, src/mutation_annotated_tree.cpp:1962-2031, is never run)
class flow_control { public: void stop() {} };
namespace flow {
class graph { public: void wait_for_all() {} };
enum { unlimited = 0, serial = 1 };
template <typename In, typename Out>
class function_node { public: template <typename B> function_node(graph&, int, B) {} };
template <typename T>
class input_node { public: template <typename B> input_node(graph&, B) {} void activate() {} };
template <typename A, typename B> void make_edge(A&, B&) {}
}  // namespace flow

}  // namespace tbb

// TEST INFRASTRUCTURE: std-backed stand-ins for the small Boost subset the reference's sources
// use, so that the UNMODIFIED reference (core/taxator.cpp + core/src/*.cpp) compiles in an image
// without Boost headers.  Nothing here contributes arithmetic to the RPA path.  Semantics that the
// reference relies on are kept deliberately:
//   * tuple<uint,int> constructed from make_tuple(uint,float) truncates the float (hh:592,661)
//   * lexical_cast<std::string>(std::string) is the identity (TaxonID is std::string)
//   * condition::wait(lock, pred) is an interruption point only while it actually blocks
#pragma once
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <regex>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <typeinfo>
#include <vector>
#include <sys/stat.h>

#define BOOST_THROW_EXCEPTION(x) throw x

namespace boost {

// ------------------------------------------------------------------------------------- tuple
template <class... T>
struct tuple : std::tuple<T...> {
  typedef std::tuple<T...> base;
  tuple() : base() {}
  tuple(const T&... a) : base(a...) {}
  template <class... U, class = typename std::enable_if<sizeof...(U) == sizeof...(T)>::type>
  tuple(const std::tuple<U...>& o) : base(convert(o, std::index_sequence_for<T...>())) {}
  template <class... U, class = typename std::enable_if<sizeof...(U) == sizeof...(T)>::type>
  tuple& operator=(const std::tuple<U...>& o) { static_cast<base&>(*this) = convert(o, std::index_sequence_for<T...>()); return *this; }
  template <int N> typename std::tuple_element<N, base>::type& get() { return std::get<N>(static_cast<base&>(*this)); }
  template <int N> const typename std::tuple_element<N, base>::type& get() const { return std::get<N>(static_cast<const base&>(*this)); }
 private:
  template <class... U, size_t... I>
  static base convert(const std::tuple<U...>& o, std::index_sequence<I...>) {
    return base(static_cast<typename std::tuple_element<I, base>::type>(std::get<I>(o))...);
  }
};
template <class... T> tuple<typename std::decay<T>::type...> make_tuple(T&&... a) { return tuple<typename std::decay<T>::type...>(a...); }
template <class... T> std::tuple<T&...> tie(T&... a) { return std::tuple<T&...>(a...); }
template <int N, class... T> typename std::tuple_element<N, std::tuple<T...>>::type& get(tuple<T...>& t) { return std::get<N>(static_cast<std::tuple<T...>&>(t)); }
template <int N, class... T> const typename std::tuple_element<N, std::tuple<T...>>::type& get(const tuple<T...>& t) { return std::get<N>(static_cast<const std::tuple<T...>&>(t)); }

// ------------------------------------------------------------------------------ lexical_cast
struct bad_lexical_cast : std::bad_cast { const char* what() const noexcept override { return "bad lexical cast"; } };
namespace detail {
template <class T, class S> struct lcast {
  static T run(const S& s) {
    std::stringstream ss;
    ss << s;
    T v;
    if (!(ss >> v)) throw bad_lexical_cast();
    ss >> std::ws;
    if (!ss.eof()) throw bad_lexical_cast();
    return v;
  }
};
template <class S> struct lcast<std::string, S> {
  static std::string run(const S& s) { std::ostringstream ss; ss << s; return ss.str(); }
};
template <> struct lcast<std::string, std::string> { static std::string run(const std::string& s) { return s; } };
template <class T> struct lcast_str {
  static T run(const std::string& s) {
    if (s.empty()) throw bad_lexical_cast();
    if (std::is_unsigned<T>::value && s[0] == '-') throw bad_lexical_cast();
    std::istringstream ss(s);
    ss.unsetf(std::ios::skipws);
    T v;
    if (!(ss >> v)) throw bad_lexical_cast();
    if (ss.peek() != std::char_traits<char>::eof()) throw bad_lexical_cast();
    return v;
  }
};
template <> struct lcast<unsigned int, std::string> : lcast_str<unsigned int> {};
template <> struct lcast<unsigned long, std::string> : lcast_str<unsigned long> {};
template <> struct lcast<int, std::string> : lcast_str<int> {};
template <> struct lcast<long, std::string> : lcast_str<long> {};
template <> struct lcast<float, std::string> : lcast_str<float> {};
template <> struct lcast<double, std::string> : lcast_str<double> {};
template <> struct lcast<unsigned short, std::string> : lcast_str<unsigned short> {};
}  // namespace detail
template <class T, class S> T lexical_cast(const S& s) { return detail::lcast<T, S>::run(s); }
template <class T> T lexical_cast(const char* s) { return detail::lcast<T, std::string>::run(std::string(s)); }
template <class T> T lexical_cast(char* s) { return detail::lcast<T, std::string>::run(std::string(s)); }

// ------------------------------------------------------------------------------------ format
class format {
 public:
  explicit format(const std::string& f) : fmt_(f) {}
  template <class T> format& operator%(const T& v) { std::ostringstream ss; ss << v; args_.push_back(ss.str()); return *this; }
  std::string str() const {
    std::string out; size_t a = 0;
    for (size_t i = 0; i < fmt_.size(); ++i) {
      if (fmt_[i] == '%' && i + 1 < fmt_.size()) {
        if (fmt_[i + 1] == '%') { out += '%'; ++i; continue; }
        size_t j = i + 1;
        while (j < fmt_.size() && !std::isalpha((unsigned char)fmt_[j]) && fmt_[j] != '%') ++j;  // flags/width/positional
        if (j < fmt_.size() && fmt_[j] == '%') { /* %N% positional */ }
        if (a < args_.size()) out += args_[a++];
        i = j;
        continue;
      }
      out += fmt_[i];
    }
    return out;
  }
  void clear() { args_.clear(); }
 private:
  std::string fmt_;
  mutable std::vector<std::string> args_;
  friend std::string str(const format&);
};
inline std::string str(const format& f) { std::string s = f.str(); f.args_.clear(); return s; }
inline std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.str(); }

// --------------------------------------------------------------------------------- smart ptr
template <class T>
class scoped_ptr {
 public:
  explicit scoped_ptr(T* p = nullptr) : p_(p) {}
  ~scoped_ptr() { delete p_; }
  scoped_ptr(const scoped_ptr&) = delete;
  scoped_ptr& operator=(const scoped_ptr&) = delete;
  void reset(T* p = nullptr) { if (p != p_) { delete p_; p_ = p; } }
  T* get() const { return p_; }
  T& operator*() const { return *p_; }
  T* operator->() const { return p_; }
  explicit operator bool() const { return p_ != nullptr; }
  bool operator!() const { return p_ == nullptr; }
 private:
  T* p_;
};

template <class T>
class ptr_vector {
 public:
  typedef size_t size_type;
  ptr_vector() {}
  explicit ptr_vector(size_t reserve) { v_.reserve(reserve); }
  ~ptr_vector() { for (T* p : v_) delete p; }
  ptr_vector(const ptr_vector&) = delete;
  void push_back(T* p) { v_.push_back(p); }
  T& operator[](size_t i) { return *v_[i]; }
  const T& operator[](size_t i) const { return *v_[i]; }
  size_t size() const { return v_.size(); }
 private:
  std::vector<T*> v_;
};

// owning list of heap objects (binner.cpp): iterators dereference to T&, a copy clones every element
template <class T>
class ptr_list {
  typedef std::list<T*> L;
 public:
  template <class It, class Ref, class Ptr>
  class iter {
   public:
    typedef std::bidirectional_iterator_tag iterator_category;
    typedef T value_type; typedef std::ptrdiff_t difference_type; typedef Ptr pointer; typedef Ref reference;
    iter() {}
    iter(It i) : i_(i) {}
    template <class It2, class R2, class P2> iter(const iter<It2, R2, P2>& o) : i_(o.base()) {}
    Ref operator*() const { return **i_; }
    Ptr operator->() const { return *i_; }
    iter& operator++() { ++i_; return *this; }
    iter operator++(int) { iter t(*this); ++i_; return t; }
    iter& operator--() { --i_; return *this; }
    bool operator==(const iter& o) const { return i_ == o.i_; }
    bool operator!=(const iter& o) const { return i_ != o.i_; }
    It base() const { return i_; }
   private:
    It i_;
  };
  typedef iter<typename L::iterator, T&, T*> iterator;
  typedef iter<typename L::const_iterator, const T&, const T*> const_iterator;
  typedef size_t size_type;
  ptr_list() {}
  ptr_list(const ptr_list& o) { for (T* p : o.l_) l_.push_back(new T(*p)); }
  ptr_list& operator=(const ptr_list&) = delete;
  ~ptr_list() { for (T* p : l_) delete p; }
  void push_back(T* p) { l_.push_back(p); }
  iterator begin() { return iterator(l_.begin()); }
  iterator end() { return iterator(l_.end()); }
  const_iterator begin() const { return const_iterator(l_.begin()); }
  const_iterator end() const { return const_iterator(l_.end()); }
  iterator erase(iterator it) { delete *it.base(); return iterator(l_.erase(it.base())); }
  bool empty() const { return l_.empty(); }
  size_type size() const { return l_.size(); }
  T& front() { return *l_.front(); }
  const T& front() const { return *l_.front(); }
  T& back() { return *l_.back(); }
 private:
  L l_;
};

template <class T>
class circular_buffer {
 public:
  typedef size_t size_type;
  typedef T value_type;
  explicit circular_buffer(size_type cap) : cap_(cap) {}
  void push_front(const T& v) { d_.push_front(v); if (d_.size() > cap_) d_.pop_back(); }
  T& operator[](size_type i) { return d_[i]; }
  size_type size() const { return d_.size(); }
  size_type capacity() const { return cap_; }
 private:
  size_type cap_;
  std::deque<T> d_;
};

// -------------------------------------------------------------------------------- filesystem
namespace filesystem {
inline bool exists(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
}

// ------------------------------------------------------------------------------------- regex
class regex : public std::regex {   // + size() of the expression text (binner.cpp:47)
 public:
  regex() {}
  regex(const std::string& s) : std::regex(s), n_(s.size()) {}
  regex(const char* s) : std::regex(s), n_(std::string(s).size()) {}
  size_t size() const { return n_; }
 private:
  size_t n_ = 0;
};
typedef std::cmatch cmatch;
typedef std::regex_error regex_error;
using std::regex_match;

namespace math { template <class T> bool isnan(T v) { return std::isnan(v); } }

// --------------------------------------------------------------------------------- exception
template <class Tag, class T>
struct error_info {
  typedef T value_type;
  error_info(const T& v) : value(v) {}
  typename std::remove_const<T>::type value;
};
class exception {
 public:
  virtual ~exception() {}
  mutable std::map<std::string, std::shared_ptr<void>> info_;
  mutable std::vector<std::string> text_;
};
namespace detail { template <class Tag> const char* tag_name() { return __PRETTY_FUNCTION__; } }
template <class E, class Tag, class T>
typename std::enable_if<std::is_base_of<exception, E>::value, const E&>::type operator<<(const E& e, const error_info<Tag, T>& v) {
  typedef typename std::remove_const<T>::type V;
  e.info_[detail::tag_name<Tag>()] = std::make_shared<V>(v.value);
  std::ostringstream ss; ss << v.value;
  e.text_.push_back(ss.str());
  return e;
}
template <class Info, class E>
const typename std::remove_const<typename Info::value_type>::type* get_error_info(const E& e) {
  return nullptr;  // only used by taxknife's filter (not built here)
}
inline std::string diagnostic_information(const exception& e) {
  std::string s;
  for (const auto& t : e.text_) { s += t; s += '\n'; }
  return s;
}

// ------------------------------------------------------------------------------------ thread
struct thread_interrupted {};
namespace detail { inline std::atomic<bool>& interrupt_flag() { static std::atomic<bool> f(false); return f; } }

class mutex {
 public:
  void lock() { m_.lock(); }
  void unlock() { m_.unlock(); }
  bool try_lock() { return m_.try_lock(); }
  std::mutex& native() { return m_; }
  class scoped_lock {
   public:
    explicit scoped_lock(mutex& m) : l_(m.m_) {}
    void unlock() { l_.unlock(); }
    void lock() { l_.lock(); }
    std::unique_lock<std::mutex>& native() { return l_; }
   private:
    std::unique_lock<std::mutex> l_;
  };
 private:
  std::mutex m_;
};
template <class M>
class unique_lock {
 public:
  explicit unique_lock(M& m) : l_(m.native()) {}
  void unlock() { l_.unlock(); }
  std::unique_lock<std::mutex>& native() { return l_; }
 private:
  std::unique_lock<std::mutex> l_;
};
class condition {
 public:
  template <class Lock, class Pred>
  void wait(Lock& lock, Pred pred) {
    while (!pred()) {
      if (detail::interrupt_flag().load()) throw thread_interrupted();
      cv_.wait_for(lock.native(), std::chrono::milliseconds(5));
    }
  }
  template <class Lock>
  void wait(Lock& lock) { cv_.wait_for(lock.native(), std::chrono::milliseconds(20)); }  // may wake spuriously
  void notify_one() { cv_.notify_one(); }
  void notify_all() { cv_.notify_all(); }
 private:
  std::condition_variable cv_;
};
class thread {
 public:
  static unsigned hardware_concurrency() { return std::thread::hardware_concurrency(); }
};
class thread_group {
 public:
  ~thread_group() { join_all(); }
  template <class F> void create_thread(F f) { t_.emplace_back(f); }
  void interrupt_all() { detail::interrupt_flag().store(true); }
  void join_all() { for (auto& t : t_) if (t.joinable()) t.join(); }
 private:
  std::vector<std::thread> t_;
};
using std::bind;
using std::ref;
using std::cref;

// --------------------------------------------------------------------------- string algorithms
enum token_compress_mode_type { token_compress_on, token_compress_off };
struct is_any_of { std::string set; explicit is_any_of(const std::string& s) : set(s) {} bool operator()(char c) const { return set.find(c) != std::string::npos; } };
inline bool starts_with(const std::string& s, const std::string& p) { return s.compare(0, p.size(), p) == 0; }
template <class C, class P>
C& split(C& out, const std::string& s, P pred, token_compress_mode_type mode = token_compress_off) {
  out.clear();
  std::string cur;
  bool last_sep = false;
  for (char c : s) {
    if (pred(c)) {
      if (!(mode == token_compress_on && last_sep)) { out.push_back(cur); cur.clear(); }
      last_sep = true;
    } else { cur += c; last_sep = false; }
  }
  out.push_back(cur);
  return out;
}

// --------------------------------------------------------------------------- program_options
namespace program_options {
class value_semantic {
 public:
  virtual ~value_semantic() {}
  virtual void parse(const std::vector<std::string>& tokens) = 0;
  virtual void apply_default() = 0;
  virtual bool has_default() const = 0;
  virtual bool is_multitoken() const = 0;
  virtual bool is_bool() const { return false; }
};
namespace detail {
template <class T> struct parse_one { static T run(const std::string& s) { return boost::lexical_cast<T>(s); } };
template <> struct parse_one<std::string> { static std::string run(const std::string& s) { return s; } };
template <> struct parse_one<bool> {
  static bool run(const std::string& s) {
    if (s == "1" || s == "true" || s == "on" || s == "yes") return true;
    if (s == "0" || s == "false" || s == "off" || s == "no") return false;
    throw std::runtime_error("bad bool option value: " + s);
  }
};
}  // namespace detail
template <class T>
class typed_value : public value_semantic {
 public:
  explicit typed_value(T* store) : store_(store) {}
  typed_value* default_value(const T& v) { def_ = v; has_def_ = true; return this; }
  typed_value* multitoken() { return this; }
  typed_value* required() { return this; }
  void parse(const std::vector<std::string>& t) override { if (t.empty()) throw std::runtime_error("option needs a value"); if (store_) *store_ = detail::parse_one<T>::run(t[0]); }
  void apply_default() override { if (has_def_ && store_) *store_ = def_; }
  bool has_default() const override { return has_def_; }
  bool is_multitoken() const override { return false; }
 private:
  T* store_; T def_{}; bool has_def_ = false;
};
template <class E>
class typed_value<std::vector<E>> : public value_semantic {
 public:
  explicit typed_value(std::vector<E>* store) : store_(store) {}
  typed_value* default_value(const std::vector<E>& v) { def_ = v; has_def_ = true; return this; }
  typed_value* multitoken() { multi_ = true; return this; }
  typed_value* required() { return this; }
  void parse(const std::vector<std::string>& t) override { if (store_) for (const auto& s : t) store_->push_back(detail::parse_one<E>::run(s)); }
  void apply_default() override { if (has_def_ && store_) *store_ = def_; }
  bool has_default() const override { return has_def_; }
  bool is_multitoken() const override { return multi_; }
 private:
  std::vector<E>* store_; std::vector<E> def_; bool has_def_ = false; bool multi_ = false;
};
template <class T> typed_value<T>* value(T* store = nullptr) { return new typed_value<T>(store); }

struct option_desc { std::string lng; char shrt = 0; std::shared_ptr<value_semantic> sem; std::string help; };

class options_description;
class options_easy_init {
 public:
  explicit options_easy_init(options_description* o) : o_(o) {}
  options_easy_init& operator()(const char* name, const char* help);
  options_easy_init& operator()(const char* name, value_semantic* s, const char* help = "");
 private:
  options_description* o_;
};
class options_description {
 public:
  options_description() {}
  explicit options_description(const std::string& caption) : caption_(caption) {}
  options_easy_init add_options() { return options_easy_init(this); }
  options_description& add(const options_description& o) { for (const auto& d : o.opts_) opts_.push_back(d); return *this; }
  std::vector<option_desc> opts_;
  std::string caption_;
};
inline void add_named(options_description* o, const char* name, value_semantic* s, const char* help) {
  option_desc d; std::string n(name);
  size_t c = n.find(',');
  if (c != std::string::npos) { d.lng = n.substr(0, c); d.shrt = n[c + 1]; } else d.lng = n;
  d.sem.reset(s); d.help = help ? help : "";
  o->opts_.push_back(d);
}
inline options_easy_init& options_easy_init::operator()(const char* name, const char* help) { add_named(o_, name, nullptr, help); return *this; }
inline options_easy_init& options_easy_init::operator()(const char* name, value_semantic* s, const char* help) { add_named(o_, name, s, help); return *this; }
inline std::ostream& operator<<(std::ostream& os, const options_description& o) {
  os << o.caption_ << ":\n";
  for (const auto& d : o.opts_) {
    os << "  ";
    if (d.shrt) os << '-' << d.shrt << " [ --" << d.lng << " ]"; else os << "--" << d.lng;
    if (d.sem) os << " arg";
    os << "\t" << d.help << "\n";
  }
  return os;
}

struct parsed_options { const options_description* desc; std::vector<std::pair<std::string, std::vector<std::string>>> items; };
struct variable_value {   // vm["key"].as< const std::vector<std::string> >(): the raw tokens of the option
  std::vector<std::string> toks;
  template <class T> const std::vector<std::string>& as() const { return toks; }
};
class variables_map {
 public:
  size_t count(const std::string& k) const { return seen_.count(k) ? 1 : 0; }
  const variable_value& operator[](const std::string& k) const { return raw_.at(k); }
  std::map<std::string, int> seen_;
  std::map<std::string, variable_value> raw_;
};
class command_line_parser {
 public:
  command_line_parser(int argc, char** argv) { for (int i = 1; i < argc; ++i) args_.push_back(argv[i]); }
  command_line_parser& options(const options_description& d) { desc_ = &d; return *this; }
  parsed_options run() {
    parsed_options po; po.desc = desc_;
    for (size_t i = 0; i < args_.size(); ++i) {
      const std::string& a = args_[i];
      const option_desc* od = nullptr; std::string inline_val; bool has_inline = false;
      if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
        std::string n = a.substr(2); size_t eq = n.find('=');
        if (eq != std::string::npos) { inline_val = n.substr(eq + 1); has_inline = true; n = n.substr(0, eq); }
        for (const auto& d : desc_->opts_) if (d.lng == n) od = &d;
      } else if (a.size() >= 2 && a[0] == '-') {
        for (const auto& d : desc_->opts_) if (d.shrt == a[1]) od = &d;
        if (a.size() > 2) { inline_val = a.substr(2); has_inline = true; }
      }
      if (!od) throw std::runtime_error("unrecognised option '" + a + "'");
      std::vector<std::string> toks;
      if (od->sem) {
        if (has_inline) toks.push_back(inline_val);
        else if (od->sem->is_multitoken()) { while (i + 1 < args_.size() && args_[i + 1][0] != '-') toks.push_back(args_[++i]); }
        else if (i + 1 < args_.size()) toks.push_back(args_[++i]);
        else throw std::runtime_error("option '" + a + "' needs a value");
      }
      po.items.emplace_back(od->lng, toks);
    }
    return po;
  }
 private:
  std::vector<std::string> args_;
  const options_description* desc_ = nullptr;
};
inline void store(const parsed_options& po, variables_map& vm) {
  for (const auto& d : po.desc->opts_) if (d.sem) { d.sem->apply_default(); if (d.sem->has_default()) vm.seen_[d.lng] = 1; }
  for (const auto& it : po.items) {
    for (const auto& d : po.desc->opts_) if (d.lng == it.first) {
      if (d.sem) d.sem->parse(it.second);
      vm.seen_[d.lng] = 1;
      for (const auto& t : it.second) vm.raw_[d.lng].toks.push_back(t);
    }
  }
}
inline void notify(variables_map&) {}
}  // namespace program_options

}  // namespace boost

// Minimal stand-in for the slice of boost::program_options the reference's src/WEPP/util.cpp
// (:135-186) and src/WEPP/dataset.hpp (:11-56) touch.  variables_map is a string->any map the
// driver fills directly; the command-line parser entry points exist only so util.cpp compiles.
// TEST INFRASTRUCTURE; written from the documented Boost interface, nothing copied.
#pragma once
#include <any>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost {
namespace program_options {

class variable_value {
  public:
    variable_value() = default;
    explicit variable_value(std::any v) : v_(std::move(v)) {}
    template <typename T> const T& as() const { return *std::any_cast<T>(&v_); }
    bool empty() const { return !v_.has_value(); }
  private:
    std::any v_;
};

class variables_map : public std::map<std::string, variable_value> {
  public:
    const variable_value& operator[](const std::string& k) const {
        auto it = this->find(k);
        if (it == this->end()) throw std::runtime_error("option not set: " + k);
        return it->second;
    }
    template <typename T> void set(const std::string& k, T v) {
        std::map<std::string, variable_value>::operator[](k) = variable_value(std::any(std::move(v)));
    }
    size_t count(const std::string& k) const { return std::map<std::string, variable_value>::count(k); }
};

template <typename T>
class typed_value {
  public:
    typed_value* default_value(const T&) { return this; }
    template <typename U> typed_value* default_value(const U&) { return this; }
    typed_value* required() { return this; }
};
template <typename T>
typed_value<T>* value() { static typed_value<T> v; return &v; }

class options_description;
class options_description_easy_init {
  public:
    options_description_easy_init& operator()(const char*, const char*) { return *this; }
    template <typename V> options_description_easy_init& operator()(const char*, V*, const char*) { return *this; }
    template <typename V> options_description_easy_init& operator()(const char*, V*) { return *this; }
};
class options_description {
  public:
    options_description() = default;
    explicit options_description(const std::string&) {}
    options_description_easy_init add_options() { return {}; }
    options_description& add(const options_description&) { return *this; }
};
inline std::ostream& operator<<(std::ostream& os, const options_description&) { return os; }

struct option { std::vector<std::string> value; std::string string_key; bool unregistered = false; };
class parsed_options {
  public:
    std::vector<option> options;
};
enum collect_unrecognized_mode { include_positional, exclude_positional };
inline std::vector<std::string> collect_unrecognized(const std::vector<option>& opts, collect_unrecognized_mode) {
    std::vector<std::string> r;
    for (const auto& o : opts) for (const auto& v : o.value) r.push_back(v);
    return r;
}
class command_line_parser {
  public:
    explicit command_line_parser(const std::vector<std::string>&) {}
    command_line_parser(int, char**) {}
    command_line_parser& options(const options_description&) { return *this; }
    command_line_parser& allow_unregistered() { return *this; }
    parsed_options run() { return {}; }
};
inline void store(const parsed_options&, variables_map&) {}
inline void notify(variables_map&) {}

}  // namespace program_options
}  // namespace boost

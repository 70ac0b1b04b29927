/*
 * ParameterFile.hpp — the parameter-file surface of the reference, re-implemented
 * for the B200 host layer.  A user's existing `.param` files must mean exactly the
 * same thing here.
 *
 * Mirrors (behaviour, not code):
 *   YAML-subset dictionary   /root/reference/src/YAMLDictionary.hpp:177-260 (parse),
 *                            :283-358 (used-values dump), :395-520 (typed access)
 *   ParameterFile            /root/reference/src/ParameterFile.hpp:93-155, ParameterFile.cpp:36-56
 *   unit handling            /root/reference/src/UnitConverter.hpp:103-170 (unit table),
 *                            :259-342 (energy<->frequency, wavelength<->frequency),
 *                            :362-438 (composed unit strings), Unit.hpp:84-150
 *   value parsing            /root/reference/src/Utilities.hpp:96-111 (vectors), :232-240,
 *                            :487-515 (booleans), :696-711 (value + unit)
 *
 * Format: `Group:` lines open an indentation-scoped group, `key: value` lines are
 * stored under "Group:Sub:key"; `#` starts a comment; values are strings that are
 * converted on access; physical values are "<number> <unit>" and are converted to
 * SI; every accessed key (with the default that was used) is remembered so that
 * `<file>.used-values` can be written like the reference does.
 */
#pragma once
#include <algorithm>
#include <array>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <istream>
#include <map>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

#include "Error.hpp"

namespace cmi {

/* CODATA 2014 constants used by the reference (src/PhysicalConstants.hpp:73-110) */
namespace constants {
constexpr double planck = 6.626070040e-34;
constexpr double boltzmann = 1.38064852e-23;
constexpr double lightspeed = 299792458.;
constexpr double electronvolt = 1.6021766208e-19;
constexpr double proton_mass = 1.672621898e-27;
constexpr double newton_constant = 6.67408e-11;
} // namespace constants

enum Quantity {
  QUANTITY_ACCELERATION, QUANTITY_ANGLE, QUANTITY_DENSITY, QUANTITY_ENERGY, QUANTITY_FLUX,
  QUANTITY_FREQUENCY, QUANTITY_LENGTH, QUANTITY_MASS, QUANTITY_NUMBER_DENSITY, QUANTITY_REACTION_RATE,
  QUANTITY_SURFACE_AREA, QUANTITY_TEMPERATURE, QUANTITY_TIME, QUANTITY_VELOCITY, QUANTITY_VOLUME,
  QUANTITY_SURFACE_DENSITY, /* appended: the integer values are part of the cmih_paramfile_get_physical probe */
  QUANTITY_MASS_RATE, QUANTITY_FREQUENCY_PER_MASS
};

/* a unit = SI value of one unit + exponents of (length, time, mass, temperature, angle) */
struct Unit {
  double value;
  int dim[5];
  bool same_quantity(const Unit &o) const {
    for (int k = 0; k < 5; ++k)
      if (dim[k] != o.dim[k]) return false;
    return true;
  }
  /* integer power by repeated multiplication / division, as Unit::operator^= does
   * (bit-identical conversion factors for e.g. cm^-3, cm^2) */
  void raise(int power) {
    const double base = value;
    if (power >= 0) {
      for (int i = 1; i < power; ++i) value *= base;
    } else {
      value = 1.;
      for (int i = 0; i < -power; ++i) value /= base;
    }
    for (int k = 0; k < 5; ++k) dim[k] *= power;
  }
  void multiply(const Unit &o) {
    value *= o.value;
    for (int k = 0; k < 5; ++k) dim[k] += o.dim[k];
  }
};

class UnitConverter {
public:
  static Unit single_unit(const std::string &name) {
    struct Row { const char *name; double v; int l, t, m, K, a; };
    static const Row table[] = {
        {"m", 1., 1, 0, 0, 0, 0}, {"cm", 0.01, 1, 0, 0, 0, 0}, {"pc", 3.086e16, 1, 0, 0, 0, 0},
        {"kpc", 3.086e19, 1, 0, 0, 0, 0}, {"angstrom", 1.e-10, 1, 0, 0, 0, 0}, {"km", 1000., 1, 0, 0, 0, 0},
        {"au", 149597870700., 1, 0, 0, 0, 0},
        {"s", 1., 0, 1, 0, 0, 0}, {"Gyr", 3.154e16, 0, 1, 0, 0, 0}, {"Myr", 3.154e13, 0, 1, 0, 0, 0},
        {"yr", 3.154e7, 0, 1, 0, 0, 0}, {"h", 3600., 0, 1, 0, 0, 0},
        {"kg", 1., 0, 0, 1, 0, 0}, {"g", 0.001, 0, 0, 1, 0, 0}, {"Msol", 1.98855e30, 0, 0, 1, 0, 0},
        {"K", 1., 0, 0, 0, 1, 0},
        {"radians", 1., 0, 0, 0, 0, 1}, {"degrees", M_PI / 180., 0, 0, 0, 0, 1},
        {"Hz", 1., 0, -1, 0, 0, 0},
        {"J", 1., 2, -2, 1, 0, 0}, {"erg", 1.e-7, 2, -2, 1, 0, 0}, {"eV", constants::electronvolt, 2, -2, 1, 0, 0},
        {"Pa", 1., -1, -2, 1, 0, 0}, {"bar", 1.e5, -1, -2, 1, 0, 0}};
    for (const Row &r : table)
      if (name == r.name) return Unit{r.v, {r.l, r.t, r.m, r.K, r.a}};
    cmi_error("Unknown unit: \"%s\"!", name.c_str());
  }

  static const char *si_unit_name(Quantity q) {
    switch (q) {
    case QUANTITY_ACCELERATION: return "m s^-2";
    case QUANTITY_ANGLE: return "radians";
    case QUANTITY_DENSITY: return "kg m^-3";
    case QUANTITY_ENERGY: return "J";
    case QUANTITY_FLUX: return "m^-2 s^-1";
    case QUANTITY_FREQUENCY: return "Hz";
    case QUANTITY_LENGTH: return "m";
    case QUANTITY_MASS: return "kg";
    case QUANTITY_NUMBER_DENSITY: return "m^-3";
    case QUANTITY_REACTION_RATE: return "m^3 s^-1";
    case QUANTITY_SURFACE_AREA: return "m^2";
    case QUANTITY_SURFACE_DENSITY: return "kg m^-2";
    case QUANTITY_MASS_RATE: return "kg s^-1";
    case QUANTITY_FREQUENCY_PER_MASS: return "Hz kg^-1";
    case QUANTITY_TEMPERATURE: return "K";
    case QUANTITY_TIME: return "s";
    case QUANTITY_VELOCITY: return "m s^-1";
    case QUANTITY_VOLUME: return "m^3";
    }
    cmi_error("Unknown quantity: %i!", (int)q);
  }

  /* "K kg^3 s^-1m ": names separated by blanks, optional ^power directly after a name */
  static Unit parse(const std::string &text) {
    Unit result{1., {0, 0, 0, 0, 0}};
    bool any = false;
    size_t pos = 0;
    const size_t n = text.size();
    while (pos < n) {
      while (pos < n && !isalpha((unsigned char)text[pos])) ++pos;
      if (pos == n) break;
      size_t end = pos + 1;
      while (end < n && text[end] != ' ' && text[end] != '^') ++end;
      Unit u = single_unit(text.substr(pos, end - pos));
      pos = end;
      if (pos < n && text[pos] == '^') {
        size_t p0 = ++pos;
        ++pos;
        while (pos < n && (isdigit((unsigned char)text[pos]) || text[pos] == '+' || text[pos] == '-')) ++pos;
        u.raise(std::stoi(text.substr(p0, pos - p0)));
      }
      if (any) {
        result.multiply(u);
      } else {
        result = u;
        any = true;
      }
    }
    if (!any) cmi_error("Empty unit provided!");
    return result;
  }

  static double to_SI(Quantity q, double value, const std::string &unit) {
    const Unit si = parse(si_unit_name(q));
    const Unit from = parse(unit);
    if (si.same_quantity(from)) return value * from.value;
    return convert_quantity(value, from, si);
  }

  static double convert(double value, const std::string &unit_from, const std::string &unit_to) {
    const Unit from = parse(unit_from), to = parse(unit_to);
    if (from.same_quantity(to)) return value * from.value / to.value;
    return convert_quantity(value, from, to);
  }

private:
  /* photon energy <-> frequency (x 1/h) and wavelength <-> frequency (c / x), same operation
   * order as UnitConverter::try_conversion so that e.g. "13.6 eV" is the same double */
  static double convert_quantity(double value, const Unit &from, const Unit &to) {
    const Unit energy = parse("J"), frequency = parse("Hz"), length = parse("m");
    const double inv_h = 1. / constants::planck;
    if (from.same_quantity(energy) && to.same_quantity(frequency)) return value * from.value * inv_h / to.value;
    if (from.same_quantity(frequency) && to.same_quantity(energy)) return value * from.value / inv_h / to.value;
    if (from.same_quantity(length) && to.same_quantity(frequency)) {
      const double s = value * from.value;
      return (1. / s) * constants::lightspeed / to.value;
    }
    if (from.same_quantity(frequency) && to.same_quantity(length)) {
      const double s = value * from.value / constants::lightspeed;
      return (1. / s) / to.value;
    }
    cmi_error("No known conversion between the given units!");
  }
};

namespace detail {

inline std::string strip(const std::string &s) {
  const size_t a = s.find_first_not_of(" \t");
  if (a == std::string::npos) return "";
  const size_t b = s.find_last_not_of(" \t");
  return s.substr(a, b - a + 1);
}

inline std::array<std::string, 3> split_vector(const std::string &value) {
  std::array<std::string, 3> out;
  size_t p1 = value.find('[') + 1;
  size_t p2 = value.find(',', p1);
  out[0] = value.substr(p1, p2 - p1);
  p1 = p2 + 1;
  p2 = value.find(',', p1);
  out[1] = value.substr(p1, p2 - p1);
  p1 = p2 + 1;
  p2 = value.find(']', p1);
  out[2] = value.substr(p1, p2 - p1);
  return out;
}

inline std::pair<double, std::string> split_value_unit(const std::string &s) {
  size_t idx = 0;
  double v;
  try {
    v = std::stod(s, &idx);
  } catch (std::exception &) {
    cmi_error("Error extracting value from \"%s\" unit-value pair!", s.c_str());
  }
  while (idx < s.size() && s[idx] == ' ') ++idx;
  return {v, s.substr(idx)};
}

/* integers may be written in exponent notation ("1e8") */
template <class I> I to_integer(const std::string &s) {
  if (s.empty()) cmi_error("Cannot extract an integer from an empty string!");
  char *end = nullptr;
  const long double v = strtold(s.c_str(), &end);
  if (end == s.c_str()) cmi_error("Error converting \"%s\" to an integer value!", s.c_str());
  return (I)std::llround(v);
}

template <class T> struct Convert;
template <> struct Convert<std::string> { static std::string from(const std::string &s) { return s; } static std::string str(const std::string &v) { return v; } };
template <> struct Convert<double> {
  static double from(const std::string &s) {
    char *end = nullptr;
    const double v = strtod(s.c_str(), &end);
    if (end == s.c_str()) cmi_error("Error converting \"%s\" to a floating point value!", s.c_str());
    return v;
  }
  static std::string str(double v) { std::ostringstream o; o << v; return o.str(); }
};
template <> struct Convert<bool> {
  static bool from(const std::string &s) {
    std::string t = strip(s);
    std::transform(t.begin(), t.end(), t.begin(), ::tolower);
    if (t == "true" || t == "yes" || t == "on" || t == "y") return true;
    if (t == "false" || t == "no" || t == "off" || t == "n") return false;
    cmi_error("Error converting \"%s\" to a boolean value!", t.c_str());
  }
  static std::string str(bool v) { return v ? "true" : "false"; }
};
template <class I> struct ConvertInt {
  static I from(const std::string &s) { return to_integer<I>(s); }
  static std::string str(I v) { std::ostringstream o; o << v; return o.str(); }
};
template <> struct Convert<int32_t> : ConvertInt<int32_t> {};
template <> struct Convert<uint32_t> : ConvertInt<uint32_t> {};
template <> struct Convert<int64_t> : ConvertInt<int64_t> {};
template <> struct Convert<uint64_t> : ConvertInt<uint64_t> {};
template <class T> struct ConvertVec {
  static std::array<T, 3> from(const std::string &s) {
    const auto p = split_vector(s);
    return {Convert<T>::from(strip(p[0])), Convert<T>::from(strip(p[1])), Convert<T>::from(strip(p[2]))};
  }
  static std::string str(const std::array<T, 3> &v) {
    return "[" + Convert<T>::str(v[0]) + ", " + Convert<T>::str(v[1]) + ", " + Convert<T>::str(v[2]) + "]";
  }
};
template <> struct Convert<std::array<double, 3>> : ConvertVec<double> {};
template <> struct Convert<std::array<bool, 3>> : ConvertVec<bool> {};
template <> struct Convert<std::array<int32_t, 3>> : ConvertVec<int32_t> {};
template <> struct Convert<std::array<uint32_t, 3>> : ConvertVec<uint32_t> {};

} // namespace detail

class YAMLDictionary {
public:
  YAMLDictionary() = default;

  explicit YAMLDictionary(std::istream &stream) {
    std::string line;
    std::vector<std::string> groups;
    std::vector<size_t> levels;
    while (std::getline(stream, line)) {
      const size_t first = line.find_first_not_of(" \t");
      if (first == std::string::npos || line[first] == '#') continue;
      const size_t hash = line.find('#');
      if (hash != std::string::npos) line = line.substr(0, hash);
      const size_t colon = line.find(':');
      if (colon == std::string::npos) cmi_error("Error while parsing line \"%s\": no ':' found!", line.c_str());
      const std::string name = detail::strip(line.substr(0, colon));
      const std::string value = detail::strip(line.substr(colon + 1));
      const size_t indent = first;
      if (indent > 0) {
        if (!levels.empty()) {
          if (indent > levels.back()) {
            levels.push_back(indent);
          } else {
            while (!levels.empty() && indent < levels.back()) {
              levels.pop_back();
              if (groups.empty()) cmi_error("Line has a different indentation than expected: \"%s\"!", line.c_str());
              groups.pop_back();
            }
            if (levels.empty()) cmi_error("Line has a different indentation than expected: \"%s\"!", line.c_str());
          }
        } else {
          levels.push_back(indent);
        }
        if (levels.size() != groups.size())
          cmi_error("Line has a different indentation than expected: \"%s\"!", line.c_str());
        if (value.empty()) {
          groups.push_back(name);
        } else {
          std::string key;
          for (const std::string &g : groups) key += g + ":";
          dict_[key + name] = value;
        }
      } else {
        if (groups.size() != levels.size()) cmi_error("Wrong formatting!");
        levels.clear();
        groups.clear();
        if (value.empty()) groups.push_back(name);
        else dict_[name] = value;
      }
    }
  }

  bool has_value(const std::string &key) const {
    auto it = dict_.find(key);
    return it != dict_.end() && it->second != "default value";
  }

  void add_value(const std::string &key, const std::string &value) {
    dict_[key] = value;
    used_[key] = value;
  }

  /* mandatory value */
  template <class T> T get_value(const std::string &key) {
    const T v = detail::Convert<T>::from(raw(key));
    used_[key] = detail::Convert<T>::str(v);
    return v;
  }
  /* value with default */
  template <class T> T get_value(const std::string &key, const T &default_value) {
    const std::string s = raw(key, "");
    const T v = s.empty() ? default_value : detail::Convert<T>::from(s);
    used_[key] = detail::Convert<T>::str(v);
    return v;
  }
  std::string get_value(const std::string &key, const char *default_value) {
    return raw(key, default_value);
  }

  template <Quantity Q> double get_physical_value(const std::string &key) { return physical<Q>(key, raw(key)); }
  template <Quantity Q> double get_physical_value(const std::string &key, const std::string &default_value) {
    return physical<Q>(key, raw(key, default_value));
  }
  template <Quantity Q> std::array<double, 3> get_physical_vector(const std::string &key) {
    return physical_vector<Q>(key, raw(key));
  }
  template <Quantity Q>
  std::array<double, 3> get_physical_vector(const std::string &key, const std::string &default_value) {
    return physical_vector<Q>(key, raw(key, default_value));
  }

  /* same layout as YAMLDictionary::print_contents (groups re-created from the sorted keys) */
  void print_contents(std::ostream &out, bool used_values) const {
    std::vector<std::string> open;
    for (const auto &kv : dict_) {
      std::vector<std::string> parts;
      size_t s = 0, c;
      while ((c = kv.first.find(':', s)) != std::string::npos) {
        parts.push_back(kv.first.substr(s, c - s));
        s = c + 1;
      }
      const std::string leaf = kv.first.substr(s);
      size_t common = 0;
      while (common < open.size() && common < parts.size() && open[common] == parts[common]) ++common;
      open.resize(common);
      for (size_t j = common; j < parts.size(); ++j) {
        out << std::string(2 * j, ' ') << parts[j] << ":\n";
        open.push_back(parts[j]);
      }
      const std::string indent(2 * parts.size(), ' ');
      if (used_values) {
        auto u = used_.find(kv.first);
        out << indent << leaf << ": " << (u != used_.end() ? u->second : std::string("value not used")) << " # ("
            << kv.second << ")\n";
      } else {
        out << indent << leaf << ": " << kv.second << "\n";
      }
    }
  }

  const std::map<std::string, std::string> &used_values() const { return used_; }

private:
  std::map<std::string, std::string> dict_, used_;

  std::string raw(const std::string &key) {
    auto it = dict_.find(key);
    if (it == dict_.end()) cmi_error("Parameter \"%s\" not found!", key.c_str());
    used_[key] = it->second;
    return it->second;
  }
  std::string raw(const std::string &key, const std::string &default_value) {
    auto it = dict_.find(key);
    std::string s;
    if (it == dict_.end() || it->second == "default value") {
      dict_[key] = "default value";
      s = default_value;
    } else {
      s = it->second;
    }
    used_[key] = s;
    return s;
  }
  template <Quantity Q> double physical(const std::string &key, const std::string &s) {
    const auto vu = detail::split_value_unit(s);
    const double v = UnitConverter::to_SI(Q, vu.first, vu.second);
    used_[key] = detail::Convert<double>::str(v) + " " + UnitConverter::si_unit_name(Q);
    return v;
  }
  template <Quantity Q> std::array<double, 3> physical_vector(const std::string &key, const std::string &s) {
    const auto parts = detail::split_vector(s);
    std::array<double, 3> v;
    std::string used = "[";
    for (int k = 0; k < 3; ++k) {
      const auto vu = detail::split_value_unit(detail::strip(parts[k]));
      v[k] = UnitConverter::to_SI(Q, vu.first, vu.second);
      used += detail::Convert<double>::str(v[k]) + " " + UnitConverter::si_unit_name(Q);
      if (k < 2) used += ", ";
    }
    used_[key] = used + "]";
    return v;
  }
};

class ParameterFile : public YAMLDictionary {
public:
  ParameterFile() = default;
  explicit ParameterFile(const std::string &filename) : filename_(filename) {
    std::ifstream file(filename);
    if (!file) cmi_error("Failed to open parameter file \"%s\"", filename.c_str());
    static_cast<YAMLDictionary &>(*this) = YAMLDictionary(file);
  }
  /* a file name given in the parameter file; relative names are relative to the working
   * directory, as in the reference (ParameterFile::get_filename) */
  std::string get_filename(const std::string &key) { return get_value<std::string>(key); }
  std::string get_filename(const std::string &key, const std::string &default_value) {
    return get_value<std::string>(key, default_value);
  }

  void print_contents(std::ostream &out) const {
    const time_t now = time(nullptr);
    char stamp[64];
    strftime(stamp, sizeof(stamp), "%d/%m/%Y, %H:%M:%S", localtime(&now));
    out << "# file written on " << stamp << ".\n";
    YAMLDictionary::print_contents(out, true);
  }
  const std::string &filename() const { return filename_; }

private:
  std::string filename_;
};

} // namespace cmi

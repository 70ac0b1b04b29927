/*
 * HostCommon.hpp — what every header of the host layer shares: the C ABI, the parameter file, logging, the ion /
 * element names of the parameter files, SimulationBox (reference: Log.hpp, ElementNames.hpp, SimulationBox.hpp).
 */
#pragma once
#include <array>
#include <cfloat>
#include <chrono>
#include <cinttypes>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <dlfcn.h>
#include <sys/utsname.h>
#include <condition_variable>
#include <mutex>
#include "../../include/cmib.h"
#include "Error.hpp"
#include "HDF5Reader.hpp"
#include "HDF5Writer.hpp"
#include "ParameterFile.hpp"
#include "RandomGenerator.hpp"
#include "../csrc/spectrum_tables.hpp" /* host-side table builders + the samplers the device uses (plain C++) */

namespace cmi {

using Vec3 = std::array<double, 3>;

/* ---- logging: same levels as the reference's Log (Log.hpp:41-46), terminal only ---- */
class Log {
public:
  enum Level { INFO = 0, STATUS, WARNING, ERROR_ };
  explicit Log(Level level = STATUS, std::ostream &out = std::cerr) : level_(level), out_(out) {}
  template <class... A> void write_info(const A &...a) { write(INFO, a...); }
  template <class... A> void write_status(const A &...a) { write(STATUS, a...); }
  template <class... A> void write_warning(const A &...a) { write(WARNING, a...); }

private:
  Level level_;
  std::ostream &out_;
  template <class... A> void write(Level l, const A &...a) {
    if (l < level_) return;
    std::ostringstream s;
    (void)std::initializer_list<int>{(s << a, 0)...};
    out_ << s.str() << "\n";
  }
};

#define CMIB_CALL(expr)                                                                         \
  do {                                                                                          \
    if ((expr) != 0) cmi_error("%s failed: %s", #expr, cmib_last_error());                      \
  } while (0)

/* ---- ion / element names of the parameter files (ElementNames.hpp:107-160, 52-88) ---- */
inline const char *ion_name(int ion) {
  static const char *names[CMIB_NUM_IONS] = {"H_n", "He_n", "C_p1", "C_p2", "N_n", "N_p1", "N_p2",
                                             "O_n", "O_p1", "Ne_n", "Ne_p1", "S_p1", "S_p2", "S_p3"};
  return names[ion];
}
/* get_ion_name (ElementNames.hpp:210-240): the names snapshot fields carry (NeutralFractionH, NeutralFractionC+, ...) */
inline const char *ion_symbol(int ion) {
  static const char *names[CMIB_NUM_IONS] = {"H", "He", "C+", "C++", "N", "N+", "N++", "O", "O+", "Ne", "Ne+", "S+", "S++", "S+++"};
  return names[ion];
}
inline const char *element_name(int el) {
  static const char *names[CMIB_NUM_ELEMENTS] = {"He", "C", "N", "O", "Ne", "S"};
  return names[el];
}

struct SimulationBox {
  Vec3 anchor, sides;
  std::array<bool, 3> periodicity;
  explicit SimulationBox(ParameterFile &params)
      : anchor(params.get_physical_vector<QUANTITY_LENGTH>("SimulationBox:anchor", "[-5. pc, -5. pc, -5. pc]")),
        sides(params.get_physical_vector<QUANTITY_LENGTH>("SimulationBox:sides", "[10. pc, 10. pc, 10. pc]")),
        periodicity(params.get_value<std::array<bool, 3>>("SimulationBox:periodicity", {false, false, false})) {}
};

} // namespace cmi

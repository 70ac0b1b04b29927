/*
 * DensityGridWriters.hpp — DensityGridWriter plugins: the Gadget-style HDF5 snapshot (the reference's default) and the ASCII layout.
 * Part of the host layer described in IonizationSimulation.hpp (class map, reference citations).
 */
#pragma once
#include "HostCommon.hpp"
#include "DensityGrid.hpp"
#include "HDF5Writer.hpp"

namespace cmi {

/* ---- writers ---- */
class DensityGridWriter {
public:
  virtual ~DensityGridWriter() {}
  /* DensityGridWriter::write(grid, iteration, params, time) (DensityGridWriter.hpp); works on the host mirror
   * of the cells: the caller refreshes it (CartesianDensityGrid::download) */
  virtual void write(CartesianCells &grid, uint32_t iteration, ParameterFile &params, double time = 0.) = 0;
};

/* the reference's ASCII snapshot layout, optionally with every field */
class AsciiFileDensityGridWriter : public DensityGridWriter {
public:
  AsciiFileDensityGridWriter(std::string prefix, std::string output_folder, bool all_fields = false)
      : prefix_(std::move(prefix)), folder_(std::move(output_folder)), all_fields_(all_fields) {}
  AsciiFileDensityGridWriter(const std::string &output_folder, ParameterFile &params)
      : AsciiFileDensityGridWriter(params.get_value<std::string>("DensityGridWriter:prefix", "snapshot"), output_folder,
                                   params.get_value<bool>("DensityGridWriter:all fields", false)) {}
  std::string filename(uint32_t iteration) const {
    char num[16];
    snprintf(num, sizeof(num), "%03u", iteration);
    return folder_ + "/" + prefix_ + num + ".txt";
  }
  void write(CartesianCells &grid, uint32_t iteration, ParameterFile &, double = 0.) override { write(grid, iteration); }
  void write(CartesianCells &grid, uint32_t iteration) {
    std::ofstream file(filename(iteration));
    if (!file) cmi_error("Unable to open snapshot file \"%s\"!", filename(iteration).c_str());
    const size_t n = grid.get_number_of_cells();
    const double volume = grid.get_cell_volume();
    if (!all_fields_) {
      /* AsciiFileDensityGridWriter.cpp:75-95 */
      file << "#x (m)\ty (m)\tz (m)\tn (m^-3)\tvolume (m^3)\tneutral H fraction\n";
      for (size_t i = 0; i < n; ++i) {
        const Vec3 x = grid.get_cell_midpoint(i);
        file << x[0] << "\t" << x[1] << "\t" << x[2] << "\t" << grid.number_density[i] << "\t" << volume << "\t"
             << grid.ionic_fraction[i] << "\n";
      }
    } else {
      file << "#x (m)\ty (m)\tz (m)\tn (m^-3)\tvolume (m^3)\tT (K)";
      for (int ion = 0; ion < CMIB_NUM_IONS; ++ion) file << "\tNeutralFraction" << ion_name(ion);
      file << "\n" << std::setprecision(17);
      for (size_t i = 0; i < n; ++i) {
        const Vec3 x = grid.get_cell_midpoint(i);
        file << x[0] << "\t" << x[1] << "\t" << x[2] << "\t" << grid.number_density[i] << "\t" << volume << "\t"
             << grid.temperature[i];
        for (int ion = 0; ion < CMIB_NUM_IONS; ++ion) file << "\t" << grid.ionic_fraction[(size_t)ion * n + i];
        file << "\n";
      }
    }
  }

private:
  std::string prefix_, folder_;
  bool all_fields_;
};

/* Which cell properties a snapshot holds: the `DensityGridWriterFields:` block
 * (DensityGridWriterFields.hpp:790-835).  Without hydro the defaults are Coordinates, NumberDensity and
 * NeutralFractionH; every `NeutralFraction<ion>` and `Temperature` can be switched on.  As in the reference a
 * flagged ion also switches on the ions before it (`ion_present` shifts the flag word, :843-846). */
struct DensityGridWriterFields {
  bool coordinates, number_density, temperature;
  uint32_t neutral_fraction = 0;
  explicit DensityGridWriterFields(ParameterFile &params) {
    coordinates = params.get_value<uint32_t>("DensityGridWriterFields:Coordinates", 1) > 0;
    number_density = params.get_value<uint32_t>("DensityGridWriterFields:NumberDensity", 1) > 0;
    temperature = params.get_value<uint32_t>("DensityGridWriterFields:Temperature", 0) > 0;
    for (int ion = 0; ion < CMIB_NUM_IONS; ++ion)
      neutral_fraction += params.get_value<uint32_t>(std::string("DensityGridWriterFields:NeutralFraction") + ion_symbol(ion),
                                                     ion == 0 ? 1u : 0u)
                          << ion;
    if (params.get_value<uint32_t>("DensityGridWriterFields:CosmicRayFactor", 0) > 0)
      cmi_error("DensityGridWriterFields:CosmicRayFactor is not provided by the B200 backend!");
  }
  bool ion_present(int ion) const { return (neutral_fraction >> ion) > 0; }
};

/* Gadget-style HDF5 snapshot, group for group and attribute for attribute what GadgetDensityGridWriter::write
 * produces (GadgetDensityGridWriter.cpp:122-300): /Header, /Code, /Configuration, /Parameters (the used values),
 * /RuntimePars, /Units (SI) and /PartType0 with Coordinates (relative to the box anchor), NumberDensity,
 * Temperature and NeutralFraction<ion>, so that the reference's benchmark analysis scripts read it
 * unchanged.  Written by host/HDF5Writer.hpp; datasets are contiguous, never compressed. */
class GadgetDensityGridWriter : public DensityGridWriter {
public:
  GadgetDensityGridWriter(std::string prefix, std::string output_folder, const DensityGridWriterFields &fields,
                          uint32_t padding = 3)
      : prefix_(std::move(prefix)), folder_(std::move(output_folder)), fields_(fields), padding_(padding) {}
  GadgetDensityGridWriter(const std::string &output_folder, ParameterFile &params)
      : GadgetDensityGridWriter(params.get_value<std::string>("DensityGridWriter:prefix", "snapshot"), output_folder,
                                DensityGridWriterFields(params), params.get_value<uint32_t>("DensityGridWriter:padding", 3)) {
    if (params.get_value<bool>("DensityGridWriter:compression", false))
      cmi_error("DensityGridWriter:compression is not provided by the B200 backend!");
  }
  /* Utilities::compose_filename: folder/prefixNNN.hdf5 */
  std::string filename(uint32_t iteration) const {
    char num[32];
    snprintf(num, sizeof(num), "%0*u", (int)padding_, iteration);
    return folder_ + "/" + prefix_ + num + ".hdf5";
  }
  void write(CartesianCells &grid, uint32_t iteration, ParameterFile &params, double time = 0.) override {
    const size_t n = grid.get_number_of_cells();
    hdf5::HDF5File file;
    hdf5::Group &header = file.root().create_group("Header");
    header.write_attribute("BoxSize", grid.get_box_sides());
    header.write_attribute("Dimension", int32_t(3));
    header.write_attribute("Flag_Entropy_ICs", std::vector<uint32_t>(6, 0));
    header.write_attribute("MassTable", std::vector<double>(6, 0.));
    header.write_attribute("NumFilesPerSnapshot", int32_t(1));
    std::vector<uint32_t> numpart(6, 0);
    numpart[0] = (uint32_t)n;
    header.write_attribute("NumPart_ThisFile", numpart);
    header.write_attribute("NumPart_Total", numpart);
    header.write_attribute("NumPart_Total_HighWord", std::vector<uint32_t>(6, 0));
    header.write_attribute("Time", time);

    hdf5::Group &code = file.root().create_group("Code");
    struct utsname os;
    if (uname(&os) != 0) memset(&os, 0, sizeof(os));
    code.write_attribute("Git version", "cmacionize_b200 (C ABI " + std::to_string(cmib_abi_version()) + ")");
    code.write_attribute("Compilation date", __DATE__);
    code.write_attribute("Compilation time", __TIME__);
    code.write_attribute("Compiler", std::string("GNU ") + __VERSION__);
    code.write_attribute("Operating system", os.sysname);
    code.write_attribute("Kernel name", std::string(os.sysname) + " " + os.release);
    code.write_attribute("Hardware name", os.machine);
    code.write_attribute("Host name", os.nodename);

    hdf5::Group &configuration = file.root().create_group("Configuration");
    configuration.write_attribute("BACKEND", "B200 (sm_100a) photoionization hot path, libcmib.so");
    configuration.write_attribute("HAVE_HDF5", "False (built-in writer: host/HDF5Writer.hpp)");
    configuration.write_attribute("NUMBER_OF_IONNAMES", std::to_string(CMIB_NUM_IONS));

    hdf5::Group &parameters = file.root().create_group("Parameters");
    for (const auto &kv : params.used_values()) parameters.write_attribute(kv.first, kv.second);

    hdf5::Group &runtime = file.root().create_group("RuntimePars");
    {
      char stamp[64];
      const time_t now = ::time(nullptr);
      struct tm tmv;
      localtime_r(&now, &tmv);
      strftime(stamp, sizeof(stamp), "%d/%m/%Y, %H:%M:%S", &tmv); /* Utilities::get_timestamp */
      runtime.write_attribute("Creation time", stamp);
    }
    runtime.write_attribute("Iteration", uint32_t(iteration));

    hdf5::Group &units = file.root().create_group("Units");
    units.write_attribute("Unit current in cgs (U_I)", 1.);
    units.write_attribute("Unit length in cgs (U_L)", 100.);
    units.write_attribute("Unit mass in cgs (U_M)", 1000.);
    units.write_attribute("Unit temperature in cgs (U_T)", 1.);
    units.write_attribute("Unit time in cgs (U_t)", 1.);

    hdf5::Group &part = file.root().create_group("PartType0");
    std::vector<double> coordinates;
    if (fields_.coordinates) {
      coordinates.resize(3 * n);
      const Vec3 &anchor = grid.get_box_anchor();
      for (size_t i = 0; i < n; ++i) {
        const Vec3 x = grid.get_cell_midpoint(i);
        for (int k = 0; k < 3; ++k) coordinates[3 * i + k] = x[k] - anchor[k];
      }
      part.create_dataset("Coordinates", hdf5::Type::F64, {n, 3}, coordinates.data());
    }
    if (fields_.number_density) part.create_dataset("NumberDensity", hdf5::Type::F64, {n}, grid.number_density.data());
    if (fields_.temperature) part.create_dataset("Temperature", hdf5::Type::F64, {n}, grid.temperature.data());
    for (int ion = 0; ion < CMIB_NUM_IONS; ++ion)
      if (fields_.ion_present(ion))
        part.create_dataset(std::string("NeutralFraction") + ion_symbol(ion), hdf5::Type::F64, {n},
                            grid.ionic_fraction.data() + (size_t)ion * n);
    file.write(filename(iteration));
  }

private:
  std::string prefix_, folder_;
  DensityGridWriterFields fields_;
  uint32_t padding_;
};

struct DensityGridWriterFactory {
  /* DensityGridWriterFactory.hpp:86-110; the default type is Gadget, as in the reference */
  static DensityGridWriter *generate(const std::string &output_folder, ParameterFile &params, Log *log = nullptr) {
    const std::string type = params.get_value<std::string>("DensityGridWriter:type", "Gadget");
    if (log) log->write_info("Requested DensityGridWriter type: ", type);
    if (type == "AsciiFile") return new AsciiFileDensityGridWriter(output_folder, params);
    if (type == "Gadget") return new GadgetDensityGridWriter(output_folder, params);
    cmi_error("Unknown DensityGridWriter type: \"%s\".", type.c_str());
  }
};

} // namespace cmi

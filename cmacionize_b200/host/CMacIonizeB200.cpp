/*
 * CMacIonizeB200 — command line driver of the B200 backend for the default
 * (photoionization) mode of the reference's executable
 * (/root/reference/src/CMacIonize.cpp:100-377, default branch :348-362):
 *
 *     CMacIonizeB200 --params <file> [--threads N] [--device D] [--gpus G] [--every-iteration-output]
 *                    [--output-statistics] [--dry-run] [--verbose] [--task-based]
 *
 * --threads is accepted for command-line compatibility and ignored.  --gpus G uses devices
 * D .. D+G-1 of this node (packets split by global id; accumulators all-reduced, update of the owned cell chunks, gather of the opacity records: include/cmib.h cmib_comm_*).  --task-based
 * (:304-338) reads the `TaskBasedIonizationSimulation:` parameter block instead of `IonizationSimulation:`
 * (host/IonizationSimulation.hpp) and runs the same GPU path.  Other modes of the reference (--rhd,
 * --dusty-radiative-transfer, --emission, --task-based-rhd) are outside the accelerated path and are rejected
 * with an error.
 */
#include <cstring>
#include <iostream>

#include "IonizationSimulation.hpp"

int main(int argc, char **argv) {
  std::string params;
  int device = 0, gpus = 1;
  bool every = false, stats = false, dry = false, verbose = false, task_based = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--params" && i + 1 < argc) params = argv[++i];
    else if (a == "--threads" && i + 1 < argc) ++i;
    else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
    else if (a == "--gpus" && i + 1 < argc) gpus = atoi(argv[++i]);
    else if (a == "--every-iteration-output") every = true;
    else if (a == "--output-statistics") stats = true;
    else if (a == "--dry-run") dry = true;
    else if (a == "--verbose") verbose = true;
    else if (a == "--task-based") task_based = true;
    else if (a == "--rhd" || a == "--dusty-radiative-transfer" || a == "--emission" || a == "--task-based-rhd") {
      std::cerr << "CMacIonizeB200: mode " << a << " is not part of the accelerated path\n";
      return 1;
    } else {
      std::cerr << "CMacIonizeB200: unknown argument " << a << "\n";
      return 1;
    }
  }
  if (params.empty()) {
    std::cerr << "usage: CMacIonizeB200 --params <parameter file> [--device D] [--gpus G] [--every-iteration-output] "
                 "[--output-statistics] [--dry-run] [--verbose]\n";
    return 1;
  }
  try {
    cmi::Log log(verbose ? cmi::Log::INFO : cmi::Log::STATUS);
    std::vector<int> devices;
    for (int g = 0; g < (gpus > 0 ? gpus : 1); ++g) devices.push_back(device + g);
    cmi::IonizationSimulation sim(true, every, stats, -1, params, devices, &log, task_based);
    if (dry) {
      log.write_warning("Dry run requested. Program will now halt.");
      return 0;
    }
    sim.initialize();
    sim.run();
    log.write_status("Program will now terminate.");
  } catch (const std::exception &e) {
    /* the reference prints the message and aborts (Error.hpp:101-106) */
    std::cerr << e.what() << std::endl;
    abort();
  }
  return 0;
}

// SOFA impulse-response sets (bbcat-dsp README:77-78: "src/SOFA.cpp | SOFA file support via the netcdf-bbc libraries";
// the sources are absent from the mounted tree, so the method names below follow the SOFA (AES69) conventions rather than
// a BBC header: parity unpinned).  Thin RAII wrapper over bbx_sofa_* (include/bbx.h): netCDF classic container only.
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "Convolver.h"

namespace bbcat {

class SOFA {
public:
  explicit SOFA(const std::string& filename) : s(0) {
    if (bbx_sofa_open(filename.c_str(), &s) != BBX_OK) throw std::runtime_error(std::string("libbbx: ") + bbx_last_error());
    bbx_sofa_get_sizes(s, &M, &R, &E, &N);
  }
  ~SOFA() { bbx_sofa_close(s); }

  uint_t get_num_measurements() const { return M; }
  uint_t get_num_receivers() const { return R; }
  uint_t get_num_emitters() const { return E; }
  uint_t get_ir_length() const { return N; }
  double get_samplerate(uint_t measurement = 0) const {
    double hz = 0.0;
    check(bbx_sofa_get_samplerate(s, measurement, &hz));
    return hz;
  }
  // impulse response (measurement, receiver, emitter) as fp32
  bool get_ir(std::vector<float>& ir, uint_t measurement, uint_t receiver, uint_t emitter = 0) const {
    ir.resize(N);
    return bbx_sofa_get_ir(s, measurement, receiver, emitter, ir.data(), N) == BBX_OK;
  }
  // Data.Delay in samples: the delay argument of Convolver::SelectFilter
  double get_delay(uint_t measurement, uint_t receiver, uint_t emitter = 0) const {
    double d = 0.0;
    check(bbx_sofa_get_delay(s, measurement, receiver, emitter, &d));
    return d;
  }
  // measurement nearest to a source position (azimuth degrees, elevation degrees, radius metres; radius <= 0: by direction)
  uint_t get_nearest_measurement(double azimuth, double elevation, double radius = 0.0) const {
    const double p[3] = {azimuth, elevation, radius};
    uint32_t m = 0;
    check(bbx_sofa_nearest_measurement(s, p, 1, &m));
    return m;
  }
  std::string get_attribute(const std::string& name) const {
    char buf[4096];
    check(bbx_sofa_get_attribute(s, name.c_str(), buf, sizeof(buf)));
    return buf;
  }
  bbx_sofa* Handle() { return s; }

private:
  static void check(int rc) {
    if (rc != BBX_OK) throw std::runtime_error(std::string("libbbx: ") + bbx_last_error());
  }
  SOFA(const SOFA&);
  SOFA& operator=(const SOFA&);
  bbx_sofa* s;
  uint32_t M, R, E, N;
};

}  // namespace bbcat

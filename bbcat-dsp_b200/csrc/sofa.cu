// sofa.cu -- bbx_sofa_*: impulse-response sets in the SOFA (AES69) conventions, the on-disk side of IR selection
// (the reference lists src/SOFA.{h,cpp} "SOFA file support via the netcdf-bbc libraries" in README:77-78 and
// libnetcdf in debian/control:5, but neither the sources nor the library are in the mounted tree or in this image).
//
// What is read: the netCDF *classic* container (CDF-1 / CDF-2, "CDF\001" / "CDF\002": big-endian header of dimension,
// attribute and variable lists followed by the variable data, record variables interleaved per record) with the SOFA
// variable names -- Data.IR [M][R][N] (FIR) or [M][R][E][N] (FIRE), Data.Delay [I|M][R]([E]), Data.SamplingRate [I|M],
// SourcePosition / ListenerPosition [I|M][C], ReceiverPosition [R][C][I|M], EmitterPosition [E][C][I|M] with their Type
// attribute -- and the global attributes (Conventions = "SOFA", DataType = "FIR" / "FIRE").  The netCDF-4 container
// (an HDF5 file, "\211HDF") that SOFA files are normally shipped in is recognised and refused with a message: reading it
// needs an HDF5 implementation, which this image does not have (`nccopy -k classic in.sofa out.nc` converts).
// Host-only code: no CUDA call here except through bbx_filter_create in bbx_sofa_create_filters.
#include <math.h>
#include <stdlib.h>

#include <map>
#include <memory>
#include <vector>

#include "common.cuh"

namespace bbx {
namespace {

enum NcType { NC_BYTE = 1, NC_CHAR = 2, NC_SHORT = 3, NC_INT = 4, NC_FLOAT = 5, NC_DOUBLE = 6 };
static uint32_t nc_size(int t) {
  switch (t) {
    case NC_BYTE:
    case NC_CHAR: return 1;
    case NC_SHORT: return 2;
    case NC_INT:
    case NC_FLOAT: return 4;
    case NC_DOUBLE: return 8;
  }
  return 0;
}

struct NcAttr {
  int type = 0;
  std::string text;             // NC_CHAR
  std::vector<double> values;   // every other type
};
struct NcVar {
  std::vector<uint32_t> dimids;
  std::map<std::string, NcAttr> attrs;
  int type = 0;
  uint64_t vsize = 0, begin = 0;
  bool record = false;
};

// bounds-checked big-endian cursor over the file image
struct Cursor {
  const uint8_t* p;
  size_t n, pos = 0;
  bool ok = true;
  Cursor(const uint8_t* d, size_t bytes) : p(d), n(bytes) {}
  bool need(size_t k) {
    if (!ok || k > n || pos > n - k) ok = false;
    return ok;
  }
  uint32_t u32() {
    if (!need(4)) return 0;
    uint32_t v = ((uint32_t)p[pos] << 24) | ((uint32_t)p[pos + 1] << 16) | ((uint32_t)p[pos + 2] << 8) | p[pos + 3];
    pos += 4;
    return v;
  }
  uint64_t u64() {
    const uint64_t hi = u32();
    return (hi << 32) | u32();
  }
  std::string name() {
    const uint32_t len = u32();
    if (!need(len)) return std::string();
    std::string s(reinterpret_cast<const char*>(p + pos), len);
    pos += len;
    skip_pad();
    return s;
  }
  void skip_pad() {
    const size_t r = pos & 3u;
    if (r) {
      if (need(4 - r)) pos += 4 - r;
    }
  }
};

static double be_value(const uint8_t* q, int type) {
  switch (type) {
    case NC_BYTE: return (double)(int8_t)q[0];
    case NC_CHAR: return (double)q[0];
    case NC_SHORT: return (double)(int16_t)(((uint16_t)q[0] << 8) | q[1]);
    case NC_INT: return (double)(int32_t)(((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3]);
    case NC_FLOAT: {
      const uint32_t u = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
      float f;
      memcpy(&f, &u, 4);
      return (double)f;
    }
    case NC_DOUBLE: {
      uint64_t u = 0;
      for (int i = 0; i < 8; i++) u = (u << 8) | q[i];
      double d;
      memcpy(&d, &u, 8);
      return d;
    }
  }
  return 0.0;
}

static const uint32_t NC_DIMENSION = 0x0A, NC_VARIABLE = 0x0B, NC_ATTRIBUTE = 0x0C;

static bool read_attrs(Cursor& c, std::map<std::string, NcAttr>& out) {
  const uint32_t tag = c.u32(), count = c.u32();
  if (!c.ok) return false;
  if (tag == 0 && count == 0) return true;  // ABSENT
  if (tag != NC_ATTRIBUTE) return false;
  for (uint32_t i = 0; i < count && c.ok; i++) {
    const std::string nm = c.name();
    NcAttr a;
    a.type = (int)c.u32();
    const uint32_t nel = c.u32(), sz = nc_size(a.type);
    if (!sz || !c.need((size_t)nel * sz)) return false;
    if (a.type == NC_CHAR) {
      a.text.assign(reinterpret_cast<const char*>(c.p + c.pos), nel);
      while (!a.text.empty() && a.text.back() == '\0') a.text.pop_back();
    } else {
      for (uint32_t k = 0; k < nel; k++) a.values.push_back(be_value(c.p + c.pos + (size_t)k * sz, a.type));
    }
    c.pos += (size_t)nel * sz;
    c.skip_pad();
    out[nm] = a;
  }
  return c.ok;
}

}  // namespace
}  // namespace bbx

using namespace bbx;

struct bbx_sofa {
  std::map<std::string, NcAttr> gattrs;
  std::vector<std::string> dim_names;
  std::vector<uint64_t> dim_len;
  // SOFA view
  uint32_t M = 0, R = 0, E = 1, N = 0;
  bool fire = false;                      // Data.IR has an emitter axis
  std::vector<float> ir;                  // [M][R][E][N]
  std::vector<double> delay;              // [Md][R][Ed] in samples, Md = 1 or M, Ed = 1 or E
  uint32_t delay_M = 1, delay_E = 1;
  std::vector<double> rate;               // [1] or [M]
  struct Pos {
    std::vector<double> v;                // [count][3]
    uint32_t count = 0;
    bool spherical = false, present = false;
  } pos[4];                               // source, listener, receiver, emitter
};

namespace bbx {
namespace {

struct NcFile {
  const uint8_t* data;
  size_t bytes;
  uint64_t numrecs = 0, recsize = 0;
  std::vector<std::string> dim_names;
  std::vector<uint64_t> dim_len;
  int rec_dim = -1;
  std::map<std::string, NcVar> vars;
  std::map<std::string, NcAttr> gattrs;
};

static int parse_header(NcFile& f) {
  BBX_REQUIRE(f.bytes >= 4, "SOFA: file shorter than a header");
  if (f.bytes >= 8 && memcmp(f.data, "\x89HDF\r\n\x1a\n", 8) == 0) {
    set_error("SOFA: this is a netCDF-4 (HDF5) container; libbbx reads the netCDF classic container only -- no HDF5 "
              "implementation is available in this build (convert with `nccopy -k classic`)");
    return BBX_ERR_UNSUPPORTED;
  }
  BBX_REQUIRE(memcmp(f.data, "CDF", 3) == 0, "SOFA: not a netCDF file (magic %02x %02x %02x %02x)", f.data[0], f.data[1], f.data[2],
              f.data[3]);
  const int version = f.data[3];
  if (version != 1 && version != 2) {
    set_error("SOFA: netCDF classic version %d is not supported (CDF-1 and CDF-2 are)", version);
    return BBX_ERR_UNSUPPORTED;
  }
  Cursor c(f.data, f.bytes);
  c.pos = 4;
  const uint32_t nrec = c.u32();
  f.numrecs = nrec == 0xFFFFFFFFu ? 0 : nrec;
  // dimensions
  {
    const uint32_t tag = c.u32(), count = c.u32();
    BBX_REQUIRE(c.ok && ((tag == 0 && count == 0) || tag == NC_DIMENSION), "SOFA: malformed dimension list");
    for (uint32_t i = 0; i < count && c.ok; i++) {
      f.dim_names.push_back(c.name());
      const uint32_t len = c.u32();
      if (len == 0) f.rec_dim = (int)i;
      f.dim_len.push_back(len);
    }
  }
  BBX_REQUIRE(read_attrs(c, f.gattrs), "SOFA: malformed global attribute list");
  // variables
  {
    const uint32_t tag = c.u32(), count = c.u32();
    BBX_REQUIRE(c.ok && ((tag == 0 && count == 0) || tag == NC_VARIABLE), "SOFA: malformed variable list");
    uint32_t nrecvars = 0;
    std::string only_rec;
    for (uint32_t i = 0; i < count && c.ok; i++) {
      const std::string nm = c.name();
      NcVar v;
      const uint32_t nd = c.u32();
      BBX_REQUIRE(c.ok && nd <= 1024, "SOFA: variable '%s' has an implausible rank", nm.c_str());
      for (uint32_t d = 0; d < nd; d++) {
        const uint32_t id = c.u32();
        BBX_REQUIRE(c.ok && id < f.dim_len.size(), "SOFA: variable '%s' refers to an unknown dimension", nm.c_str());
        v.dimids.push_back(id);
      }
      BBX_REQUIRE(read_attrs(c, v.attrs), "SOFA: malformed attribute list of '%s'", nm.c_str());
      v.type = (int)c.u32();
      v.vsize = c.u32();
      v.begin = version == 1 ? (uint64_t)c.u32() : c.u64();
      BBX_REQUIRE(c.ok && nc_size(v.type), "SOFA: variable '%s' has an unknown type", nm.c_str());
      v.record = nd > 0 && (int)v.dimids[0] == f.rec_dim;
      if (v.record) {
        f.recsize += v.vsize;
        nrecvars++;
        only_rec = nm;
      }
      f.vars[nm] = v;
    }
    BBX_REQUIRE(c.ok, "SOFA: truncated header");
    if (nrecvars == 1) {
      // a single record variable is stored without padding between its records
      const NcVar& v = f.vars[only_rec];
      uint64_t per = nc_size(v.type);
      for (size_t d = 1; d < v.dimids.size(); d++) per *= f.dim_len[v.dimids[d]];
      f.recsize = per;
    }
  }
  return BBX_OK;
}

// the variable's values in C order as doubles, its shape in `shape` (the record dimension at its current length)
static int read_var(const NcFile& f, const std::string& nm, std::vector<double>& out, std::vector<uint64_t>& shape) {
  auto it = f.vars.find(nm);
  BBX_REQUIRE(it != f.vars.end(), "SOFA: variable '%s' is missing", nm.c_str());
  const NcVar& v = it->second;
  shape.clear();
  for (size_t d = 0; d < v.dimids.size(); d++) shape.push_back(d == 0 && v.record ? f.numrecs : f.dim_len[v.dimids[d]]);
  uint64_t per = 1, total = 1;
  for (size_t d = 0; d < shape.size(); d++) {
    BBX_REQUIRE(shape[d] <= (1ull << 32), "SOFA: variable '%s' has an implausible shape", nm.c_str());
    if (!(d == 0 && v.record)) per *= shape[d];
    total *= shape[d];
    BBX_REQUIRE(total <= (1ull << 34), "SOFA: variable '%s' is too large", nm.c_str());
  }
  const uint32_t sz = nc_size(v.type);
  out.resize(total);
  const uint64_t nchunks = v.record ? f.numrecs : 1, stride = v.record ? f.recsize : 0;
  const uint64_t chunk = v.record ? per : total;
  for (uint64_t r = 0; r < nchunks; r++) {
    const uint64_t off = v.begin + r * stride;
    BBX_REQUIRE(off <= f.bytes && chunk * sz <= f.bytes - off, "SOFA: data of variable '%s' runs past the end of the file", nm.c_str());
    const uint8_t* q = f.data + off;
    for (uint64_t k = 0; k < chunk; k++) out[r * chunk + k] = be_value(q + k * sz, v.type);
  }
  return BBX_OK;
}

static bool attr_is(const std::map<std::string, NcAttr>& a, const char* key, const char* want) {
  auto it = a.find(key);
  return it != a.end() && it->second.type == NC_CHAR && it->second.text == want;
}

static uint64_t dim_by_name(const NcFile& f, const char* nm) {
  for (size_t i = 0; i < f.dim_names.size(); i++)
    if (f.dim_names[i] == nm) return (int)i == f.rec_dim ? f.numrecs : f.dim_len[i];
  return 0;
}

// [count][C] (source, listener) or [count][C][I|M] (receiver, emitter; the first measurement's geometry is kept)
static int read_position(const NcFile& f, const char* nm, bool per_object, bbx_sofa::Pos& p) {
  if (f.vars.find(nm) == f.vars.end()) return BBX_OK;  // optional
  std::vector<double> v;
  std::vector<uint64_t> sh;
  int rc = read_var(f, nm, v, sh);
  if (rc != BBX_OK) return rc;
  BBX_REQUIRE(sh.size() == (per_object ? 3u : 2u) && sh[1] == 3, "SOFA: %s must have %d dimensions with C = 3", nm, per_object ? 3 : 2);
  p.count = (uint32_t)sh[0];
  p.v.resize((size_t)p.count * 3);
  const uint64_t last = per_object ? sh[2] : 1;
  BBX_REQUIRE(last >= 1, "SOFA: %s is empty", nm);
  for (uint32_t i = 0; i < p.count; i++)
    for (int c = 0; c < 3; c++) p.v[(size_t)i * 3 + c] = v[((uint64_t)i * 3 + c) * last];
  p.spherical = attr_is(f.vars.at(nm).attrs, "Type", "spherical");
  p.present = true;
  return BBX_OK;
}

static void to_cartesian(const double* in, bool spherical, double* out) {
  if (!spherical) {
    out[0] = in[0];
    out[1] = in[1];
    out[2] = in[2];
    return;
  }
  // SOFA spherical: azimuth (degrees, counter-clockwise from +x), elevation (degrees up from the x-y plane), radius (metres)
  const double k = 3.14159265358979323846 / 180.0, az = in[0] * k, el = in[1] * k, r = in[2];
  out[0] = r * cos(el) * cos(az);
  out[1] = r * cos(el) * sin(az);
  out[2] = r * sin(el);
}

static int build(const uint8_t* data, size_t bytes, bbx_sofa** out) {
  BBX_REQUIRE(data && out, "bbx_sofa_open: null argument");
  NcFile f;
  f.data = data;
  f.bytes = bytes;
  int rc = parse_header(f);
  if (rc != BBX_OK) return rc;
  BBX_REQUIRE(attr_is(f.gattrs, "Conventions", "SOFA"), "SOFA: the global attribute Conventions is not \"SOFA\"");
  auto dt = f.gattrs.find("DataType");
  BBX_REQUIRE(dt != f.gattrs.end() && (dt->second.text == "FIR" || dt->second.text == "FIRE"),
              "SOFA: DataType '%s' is not an impulse-response set (FIR / FIRE)", dt == f.gattrs.end() ? "(missing)" : dt->second.text.c_str());
  std::unique_ptr<bbx_sofa> s(new bbx_sofa());
  s->gattrs = f.gattrs;
  s->dim_names = f.dim_names;
  s->dim_len = f.dim_len;
  if (f.rec_dim >= 0) s->dim_len[f.rec_dim] = f.numrecs;
  // Data.IR
  std::vector<double> v;
  std::vector<uint64_t> sh;
  rc = read_var(f, "Data.IR", v, sh);
  if (rc != BBX_OK) return rc;
  BBX_REQUIRE(sh.size() == 3 || sh.size() == 4, "SOFA: Data.IR must be [M][R][N] or [M][R][E][N] (rank %zu found)", sh.size());
  s->fire = sh.size() == 4;
  s->M = (uint32_t)sh[0];
  s->R = (uint32_t)sh[1];
  s->E = s->fire ? (uint32_t)sh[2] : 1;
  s->N = (uint32_t)sh.back();
  BBX_REQUIRE(s->M && s->R && s->E && s->N, "SOFA: Data.IR has an empty dimension");
  BBX_REQUIRE(dim_by_name(f, "M") == 0 || dim_by_name(f, "M") == s->M, "SOFA: Data.IR does not lead with the M dimension");
  s->ir.resize(v.size());
  for (size_t i = 0; i < v.size(); i++) s->ir[i] = (float)v[i];
  // Data.SamplingRate [I] or [M]
  rc = read_var(f, "Data.SamplingRate", s->rate, sh);
  if (rc != BBX_OK) return rc;
  BBX_REQUIRE(s->rate.size() == 1 || s->rate.size() == s->M, "SOFA: Data.SamplingRate must have 1 or M values");
  {
    const auto& a = f.vars.at("Data.SamplingRate").attrs;
    auto u = a.find("Units");
    BBX_REQUIRE(u == a.end() || u->second.text == "hertz" || u->second.text == "Hertz" || u->second.text == "Hz",
                "SOFA: Data.SamplingRate in '%s' (hertz expected)", u->second.text.c_str());
  }
  // Data.Delay [I|M][R] (FIR) or [I|M][R][E] (FIRE), in samples; optional in files written before SOFA 1.0 -> zeros
  if (f.vars.find("Data.Delay") != f.vars.end()) {
    rc = read_var(f, "Data.Delay", s->delay, sh);
    if (rc != BBX_OK) return rc;
    BBX_REQUIRE((sh.size() == 2 || sh.size() == 3) && (sh[0] == 1 || sh[0] == s->M) && sh[1] == s->R &&
                    (sh.size() == 2 || sh[2] == 1 || sh[2] == s->E),
                "SOFA: Data.Delay must be [I or M][R] or [I or M][R][E]");
    s->delay_M = (uint32_t)sh[0];
    s->delay_E = sh.size() == 3 ? (uint32_t)sh[2] : 1;
  } else {
    s->delay.assign(s->R, 0.0);
  }
  static const char* names[4] = {"SourcePosition", "ListenerPosition", "ReceiverPosition", "EmitterPosition"};
  for (int w = 0; w < 4; w++) {
    rc = read_position(f, names[w], w >= 2, s->pos[w]);
    if (rc != BBX_OK) return rc;
  }
  for (int w = 0; w < 2; w++)
    BBX_REQUIRE(!s->pos[w].present || s->pos[w].count == 1 || s->pos[w].count == s->M, "SOFA: %s must have 1 or M rows", names[w]);
  *out = s.release();
  return BBX_OK;
}

}  // namespace
}  // namespace bbx

extern "C" {

int bbx_sofa_open_memory(const void* data, size_t bytes, bbx_sofa** out) {
  return build(static_cast<const uint8_t*>(data), bytes, out);
}

int bbx_sofa_open(const char* path, bbx_sofa** out) {
  BBX_REQUIRE(path && out, "bbx_sofa_open: null argument");
  FILE* fp = fopen(path, "rb");
  BBX_REQUIRE(fp, "bbx_sofa_open: cannot open '%s'", path);
  std::vector<uint8_t> buf;
  uint8_t tmp[1 << 16];
  size_t got;
  while ((got = fread(tmp, 1, sizeof(tmp), fp)) > 0) buf.insert(buf.end(), tmp, tmp + got);
  fclose(fp);
  return build(buf.data(), buf.size(), out);
}

int bbx_sofa_close(bbx_sofa* s) {
  delete s;
  return BBX_OK;
}

int bbx_sofa_get_sizes(const bbx_sofa* s, uint32_t* M, uint32_t* R, uint32_t* E, uint32_t* N) {
  BBX_REQUIRE(s, "bbx_sofa_get_sizes: null handle");
  if (M) *M = s->M;
  if (R) *R = s->R;
  if (E) *E = s->E;
  if (N) *N = s->N;
  return BBX_OK;
}

int bbx_sofa_get_samplerate(const bbx_sofa* s, uint32_t m, double* hz) {
  BBX_REQUIRE(s && hz && m < s->M, "bbx_sofa_get_samplerate: bad argument");
  *hz = s->rate[s->rate.size() == 1 ? 0 : m];
  return BBX_OK;
}

int bbx_sofa_get_ir(const bbx_sofa* s, uint32_t m, uint32_t r, uint32_t e, float* dst, uint32_t n) {
  BBX_REQUIRE(s && dst, "bbx_sofa_get_ir: null argument");
  BBX_REQUIRE(m < s->M && r < s->R && e < s->E, "bbx_sofa_get_ir: index (%u, %u, %u) outside (%u, %u, %u)", m, r, e, s->M, s->R, s->E);
  const float* src = s->ir.data() + (((size_t)m * s->R + r) * s->E + e) * s->N;
  const uint32_t k = n < s->N ? n : s->N;
  memcpy(dst, src, (size_t)k * sizeof(float));
  if (n > k) memset(dst + k, 0, (size_t)(n - k) * sizeof(float));
  return BBX_OK;
}

int bbx_sofa_get_delay(const bbx_sofa* s, uint32_t m, uint32_t r, uint32_t e, double* samples) {
  BBX_REQUIRE(s && samples, "bbx_sofa_get_delay: null argument");
  BBX_REQUIRE(m < s->M && r < s->R && e < s->E, "bbx_sofa_get_delay: index (%u, %u, %u) outside (%u, %u, %u)", m, r, e, s->M, s->R, s->E);
  const uint32_t dm = s->delay_M == 1 ? 0 : m, de = s->delay_E == 1 ? 0 : e;
  *samples = s->delay[((size_t)dm * s->R + r) * s->delay_E + de];
  return BBX_OK;
}

int bbx_sofa_get_position(const bbx_sofa* s, int which, uint32_t index, double xyz[3], int* spherical) {
  BBX_REQUIRE(s && xyz && which >= 0 && which < 4, "bbx_sofa_get_position: bad argument");
  const bbx_sofa::Pos& p = s->pos[which];
  BBX_REQUIRE(p.present, "bbx_sofa_get_position: the file has no such variable");
  const uint32_t limit = which < 2 ? s->M : p.count;
  BBX_REQUIRE(index < limit, "bbx_sofa_get_position: index %u outside %u", index, limit);
  const uint32_t i = (which < 2 && p.count == 1) ? 0 : index;
  for (int c = 0; c < 3; c++) xyz[c] = p.v[(size_t)i * 3 + c];
  if (spherical) *spherical = p.spherical ? 1 : 0;
  return BBX_OK;
}

int bbx_sofa_nearest_measurement(const bbx_sofa* s, const double pos[3], int spherical, uint32_t* m) {
  BBX_REQUIRE(s && pos && m, "bbx_sofa_nearest_measurement: null argument");
  const bbx_sofa::Pos& p = s->pos[BBX_SOFA_SOURCE];
  BBX_REQUIRE(p.present, "bbx_sofa_nearest_measurement: the file has no SourcePosition");
  double q[3];
  // a spherical query without a usable radius selects by direction: the query becomes a unit vector and the measurements
  // are normalised below
  const bool by_direction = spherical && !(pos[2] > 0.0);
  if (by_direction) {
    const double unit[3] = {pos[0], pos[1], 1.0};
    to_cartesian(unit, true, q);
  } else {
    to_cartesian(pos, spherical != 0, q);
  }
  double best = 0.0;
  uint32_t best_m = 0;
  for (uint32_t i = 0; i < s->M; i++) {
    double c[3];
    to_cartesian(&p.v[(size_t)(p.count == 1 ? 0 : i) * 3], p.spherical, c);
    if (by_direction) {
      const double n = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
      if (n > 0.0) {
        c[0] /= n;
        c[1] /= n;
        c[2] /= n;
      }
    }
    const double d = (c[0] - q[0]) * (c[0] - q[0]) + (c[1] - q[1]) * (c[1] - q[1]) + (c[2] - q[2]) * (c[2] - q[2]);
    if (i == 0 || d < best) {  // ties keep the lowest index
      best = d;
      best_m = i;
    }
  }
  *m = best_m;
  return BBX_OK;
}

int bbx_sofa_get_attribute(const bbx_sofa* s, const char* name, char* buf, uint32_t buflen) {
  BBX_REQUIRE(s && name && buf && buflen, "bbx_sofa_get_attribute: null argument");
  auto it = s->gattrs.find(name);
  BBX_REQUIRE(it != s->gattrs.end(), "bbx_sofa_get_attribute: no global attribute '%s'", name);
  std::string text = it->second.text;
  if (it->second.type != NC_CHAR) {
    char num[64];
    text.clear();
    for (size_t i = 0; i < it->second.values.size(); i++) {
      snprintf(num, sizeof(num), i ? " %.17g" : "%.17g", it->second.values[i]);
      text += num;
    }
  }
  BBX_REQUIRE(text.size() < buflen, "bbx_sofa_get_attribute: '%s' needs %zu bytes", name, text.size() + 1);
  memcpy(buf, text.c_str(), text.size() + 1);
  return BBX_OK;
}

int bbx_sofa_create_filters(const bbx_sofa* s, bbx_engine* e, uint32_t receiver, uint32_t emitter, bbx_filter** out, uint32_t count) {
  BBX_REQUIRE(s && e && out, "bbx_sofa_create_filters: null argument");
  BBX_REQUIRE(receiver < s->R && emitter < s->E, "bbx_sofa_create_filters: receiver %u / emitter %u outside (%u, %u)", receiver, emitter,
              s->R, s->E);
  BBX_REQUIRE(count == s->M, "bbx_sofa_create_filters: room for %u filters, the set has %u measurements", count, s->M);
  for (uint32_t m = 0; m < s->M; m++) out[m] = nullptr;
  for (uint32_t m = 0; m < s->M; m++) {
    const float* h = s->ir.data() + (((size_t)m * s->R + receiver) * s->E + emitter) * s->N;
    const int rc = bbx_filter_create(e, h, s->N, &out[m]);
    if (rc != BBX_OK) {
      // a failure half way releases what was created so far (the message of the failing call is kept)
      const std::string msg = get_error();
      for (uint32_t k = 0; k < m; k++) {
        bbx_filter_destroy(out[k]);
        out[k] = nullptr;
      }
      set_error("%s", msg.c_str());
      return rc;
    }
  }
  return BBX_OK;
}

}  // extern "C"

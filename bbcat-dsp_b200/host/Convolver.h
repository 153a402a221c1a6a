// BlockConvolver / Convolver surface (bbcat-dsp README:38-44; the sources are absent from the mounted
// tree, behaviour per SURVEY.md 8.A): filter objects built from an impulse response at a partition
// size, per-channel Convolve() on fixed-size blocks, SelectFilter with crossfade and per-channel
// delay, any SampleFormat_t in and out.  Thin RAII wrappers over the bbx_engine C ABI.
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "SoundFormatConversions.h"

namespace bbcat {

class Convolver;

// Immutable filter: an IR partitioned at the convolver's block size, transformed on the GPU once.
class ConvolverFilter {
public:
  ~ConvolverFilter() { bbx_filter_destroy(f); }
  uint_t GetPartitions() const { return bbx_filter_partitions(f); }

private:
  friend class Convolver;
  explicit ConvolverFilter(bbx_filter* _f) : f(_f) {}
  ConvolverFilter(const ConvolverFilter&);
  ConvolverFilter& operator=(const ConvolverFilter&);
  bbx_filter* f;
};

// Communicator of the input-sharded MIMO convolver (one rank per GPU; NCCL is loaded by libbbx at run time).
// Rank 0 calls UniqueId() and ships the 128 bytes to the other ranks by whatever transport the application has.
class ConvolverComm {
public:
  static void UniqueId(uint8_t id[128]) {
    if (bbx_comm_unique_id(id) != BBX_OK) throw std::runtime_error(std::string("libbbx: ") + bbx_last_error());
  }
  ConvolverComm(int world, int rank, const uint8_t id[128], int device = 0) : c(0) {
    if (bbx_comm_create(world, rank, id, device, &c) != BBX_OK) throw std::runtime_error(std::string("libbbx: ") + bbx_last_error());
  }
  ~ConvolverComm() { bbx_comm_destroy(c); }
  bbx_comm* Handle() { return c; }

private:
  ConvolverComm(const ConvolverComm&);
  ConvolverComm& operator=(const ConvolverComm&);
  bbx_comm* c;
};

class Convolver {
public:
  // per-channel convolver: channel c -> filter -> delay -> channel c
  Convolver(uint_t blocksize, uint_t maxpartitions, uint_t channels, uint_t maxblocks = 1, uint_t maxdelay = 0,
            bool fractionaldelay = false, int device = 0) {
    bbx_config cfg = bbx_config();
    cfg.device = device;
    cfg.block_size = blocksize;
    cfg.max_partitions = maxpartitions;
    cfg.n_inputs = channels;
    cfg.mode = BBX_MODE_PER_CHANNEL;
    cfg.max_blocks = maxblocks;
    cfg.max_delay = maxdelay;
    cfg.fractional_delay = fractionaldelay ? 1 : 0;
    Create(cfg);
  }
  // routed (n_paths paths, time-domain mixdown) or MIMO (n_inputs x n_outputs matrix) convolver
  explicit Convolver(const bbx_config& cfg) { Create(cfg); }
  ~Convolver() { bbx_engine_destroy(e); }

  uint_t GetBlockSize() const { return blocksize; }

  ConvolverFilter* CreateFilter(const float* ir, uint_t length) {
    bbx_filter* f = 0;
    Check(bbx_filter_create(e, ir, length, &f));
    return new ConvolverFilter(f);
  }
  void SetRoute(uint_t path, uint_t input, uint_t output, float gain = 1.0f) { Check(bbx_set_route(e, path, input, output, gain)); }
  // latched; applied at the next block boundary (= first block of the next Convolve call)
  void SelectFilter(uint_t path, const ConvolverFilter* filter, double delay = 0.0, bool crossfade = false) {
    Check(bbx_set_filter(e, path, filter ? filter->f : 0, crossfade ? 1 : 0, delay));
  }
  // the same for n paths in one call (all validated before any is latched); delays / crossfade may be NULL (= 0)
  void SelectFilters(uint_t n, const uint_t* paths, const ConvolverFilter* const* filters, const double* delays = 0,
                     const int* crossfade = 0) {
    std::vector<const bbx_filter*> fv(n);
    for (uint_t k = 0; k < n; k++) fv[k] = filters[k] ? filters[k]->f : 0;
    Check(bbx_set_filters(e, n, paths, n ? &fv[0] : 0, crossfade, delays));
  }
  // nframes must be a multiple of the block size
  void Convolve(const void* src, SampleFormat_t srctype, bool src_be, uint_t src_channels, void* dst, SampleFormat_t dsttype,
                bool dst_be, uint_t dst_channels, uint_t nframes) {
    Check(bbx_process(e, src, (int)srctype, src_be, src_channels, dst, (int)dsttype, dst_be, dst_channels, nframes));
  }
  template <typename T1, typename T2>
  void Convolve(const T1* src, uint_t src_channels, T2* dst, uint_t dst_channels, uint_t nframes) {
    Convolve(src, SampleFormatOf(src), false, src_channels, dst, SampleFormatOf(dst), false, dst_channels, nframes);
  }
  // input-sharded MIMO (bbx_config::mimo_shard_world > 1): attach the communicator before the first Convolve
  void SetComm(ConvolverComm* comm) { Check(bbx_engine_set_comm(e, comm ? comm->Handle() : 0)); }
  // peer-memory mixdown of the input-sharded MIMO engine (instead of SetComm): export this rank's 64-byte handle, gather
  // all ranks' handles in rank order, attach them
  void PeerExport(uint8_t* handle64) { Check(bbx_engine_peer_export(e, handle64)); }
  void PeerAttach(const uint8_t* handles) { Check(bbx_engine_peer_attach(e, handles)); }
  bbx_engine* Handle() { return e; }

private:
  void Create(const bbx_config& cfg) {
    e = 0;
    blocksize = cfg.block_size;
    Check(bbx_engine_create(&cfg, &e));
  }
  static void Check(int rc) {
    if (rc != BBX_OK) throw std::runtime_error(std::string("libbbx: ") + bbx_last_error());
  }
  Convolver(const Convolver&);
  Convolver& operator=(const Convolver&);
  bbx_engine* e;
  uint_t blocksize;
};

// Single-channel partitioned convolution: one block in, one block out.
class BlockConvolver {
public:
  BlockConvolver(uint_t blocksize, uint_t maxpartitions, int device = 0) : conv(blocksize, maxpartitions, 1, 1, 0, false, device) {}
  ConvolverFilter* CreateFilter(const float* ir, uint_t length) { return conv.CreateFilter(ir, length); }
  void SetFilter(const ConvolverFilter* filter, bool crossfade = false) { conv.SelectFilter(0, filter, 0.0, crossfade); }
  void Convolve(const float* in, float* out) {
    if (bbx_blockconvolver_convolve(conv.Handle(), in, out) != BBX_OK) throw std::runtime_error(std::string("libbbx: ") + bbx_last_error());
  }

private:
  Convolver conv;
};

}  // namespace bbcat

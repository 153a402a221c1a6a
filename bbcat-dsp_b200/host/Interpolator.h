// Interpolator with the reference's semantics (src/Interpolator.h:12-78): `current` steps towards
// `target` by `inc`, never past it.  Layout {target, current} matches bbx_mix_samples_interp's state.
#pragma once

#include "../../include/bbx.h"

namespace bbcat {

class Interpolator {
public:
  Interpolator(float _target = 0.0f, float _current = 0.0f) {
    st[0] = _target;
    st[1] = _current;
  }
  bool NonZero() const { return (st[1] != 0.0f) || (st[0] != 0.0f); }
  Interpolator& SetCurrent(float c) { st[1] = c; return *this; }
  Interpolator& SetTarget(float t) { st[0] = t; return *this; }
  Interpolator& operator=(float t) { st[0] = t; return *this; }
  Interpolator& operator+=(float inc) { bbx_interpolator_step(st, inc, 1); return *this; }
  operator float() const { return st[1]; }
  float GetTarget() const { return st[0]; }
  bool AtTarget() const { return st[1] == st[0]; }
  float* State() { return st; }  // {target, current}

protected:
  float st[2];
};

}  // namespace bbcat

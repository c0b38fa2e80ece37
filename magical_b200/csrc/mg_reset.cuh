/*
 * mg_reset.cuh — the state half of BaseEnv.reset (magical/base_env.py:177-223) for one environment: poses from
 * the scene's reset poses, everything else zero.  Shared by k_finish / k_reset (mg_finish.cu) and
 * k_sample_layouts (mg_sample.cu).
 */
#ifndef MG_RESET_CUH
#define MG_RESET_CUH

#include "mg_device.cuh"
#include "mg_sincos.h"

#define MG_FRESH_SAMPLE 3 /* EnvState.fresh: reset from a template, the layout is still to be sampled on the device */

__device__ static __forceinline__ void mg_reset_state(EnvState& st, const DeviceScene* ds, int scene) {
  const mg_scene_t& sc = ds->s;
  st.scene = scene;
  st.episode_steps = 0;
  st.stamp = 0;
  st.n_cache = 0;
  st.overflow = 0;
  st.fresh = 1;
  st.last_contacts = 0;
  for (int b = 0; b < MG_MAX_BODIES; b++) {
    double x = 0.0, y = 0.0, a = 0.0, cs = 1.0, sn = 0.0;
    if (b < sc.n_bodies) {
      x = sc.bodies[b].p0[0]; y = sc.bodies[b].p0[1]; a = sc.bodies[b].a0;
      mg_det_sincos(a, &sn, &cs);
    }
    st.P[b] = make_double4(x, y, a, 0.0);
    st.R[b] = make_double2(cs, sn);
    st.V[b] = make_double4(0.0, 0.0, 0.0, 0.0);
    st.Bv[b] = make_double4(0.0, 0.0, 0.0, 0.0);
  }
  for (int j = 0; j < MG_MAX_JOINTS; j++) st.jacc[j] = make_double2(0.0, 0.0);
}


#endif

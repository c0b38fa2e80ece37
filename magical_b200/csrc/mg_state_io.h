/*
 * mg_state_io.h — EnvState (device record) <-> mg_state_t (the ABI's host-readable snapshot).
 * Shared by mg_api.cu (mg_get_state / mg_set_state) and the host harness of the CPU suite
 * (tests/host/tpe_host.cpp), so the conversion the GPU path uses is the one the CPU tests exercise.
 */
#ifndef MG_STATE_IO_H
#define MG_STATE_IO_H

#include <string.h>

#include "mg_device.cuh"
#include "mg_sincos.h"

static inline void mg_state_export(const EnvState& st, const mg_scene_t& sc, mg_state_t* out) {
  memset(out, 0, sizeof(*out));
  out->n_bodies = sc.n_bodies;
  out->n_joints = sc.n_joints;
  out->episode_steps = st.episode_steps;
  out->scene = st.scene;
  out->overflow = st.overflow;
  out->stamp = st.stamp;
  for (int b = 0; b < sc.n_bodies; b++) {
    out->pos[b][0] = st.P[b].x; out->pos[b][1] = st.P[b].y; out->angle[b] = st.P[b].z;
    out->vel[b][0] = st.V[b].x; out->vel[b][1] = st.V[b].y; out->angvel[b] = st.V[b].z;
    out->bias_vel[b][0] = st.Bv[b].x; out->bias_vel[b][1] = st.Bv[b].y; out->bias_angvel[b] = st.Bv[b].z;
  }
  for (int j = 0; j < sc.n_joints; j++) { out->joint_acc[j][0] = st.jacc[j].x; out->joint_acc[j][1] = st.jacc[j].y; }
  int nc = 0, ne = 0;
  for (int k = 0; k < st.n_cache && k < MG_NCACHE; k++) {
    const CEntry& e = st.cache[k];
    if (ne < MG_STATE_CACHE) {
      out->cache_shapes[ne][0] = e.a; out->cache_shapes[ne][1] = e.b;
      out->cache_hash[ne] = e.hash;
      out->cache_age[ne] = st.stamp - e.stamp;
      out->cache_jn[ne] = e.jn; out->cache_jt[ne] = e.jt;
      ne++;
    }
    if (e.stamp != st.stamp || nc >= 32) continue; /* contact_*: only the contacts of the last sub-step */
    out->contact_shapes[nc][0] = e.a; out->contact_shapes[nc][1] = e.b;
    out->contact_jn[nc] = e.jn; out->contact_jt[nc] = e.jt;
    nc++;
  }
  out->n_contacts = nc;
  out->n_cache = ne;
}

/* returns NULL on success, else why the snapshot cannot be applied to a scene with these counts */
static inline const char* mg_state_import(EnvState& st, int n_bodies, int n_joints, int n_shapes, const mg_state_t* in) {
  if (in->n_cache < 0 || in->n_cache > MG_STATE_CACHE || in->n_cache > MG_NCACHE) return "n_cache out of range";
  if (in->n_bodies != n_bodies || in->n_joints != n_joints)
    return "snapshot does not match the environment's scene (body / joint count)";
  for (int k = 0; k < in->n_cache; k++) {
    if (in->cache_shapes[k][0] < 0 || in->cache_shapes[k][0] >= n_shapes || in->cache_shapes[k][1] < 0 ||
        in->cache_shapes[k][1] >= n_shapes || in->cache_age[k] < 0 || in->cache_age[k] >= MG_PERSISTENCE)
      return "bad contact cache entry";
  }
  for (int b = 0; b < n_bodies; b++) {
    st.P[b] = make_double4(in->pos[b][0], in->pos[b][1], in->angle[b], 0.0);
    double sn, cs;
    mg_det_sincos(in->angle[b], &sn, &cs);
    st.R[b] = make_double2(cs, sn);
    st.V[b] = make_double4(in->vel[b][0], in->vel[b][1], in->angvel[b], 0.0);
    st.Bv[b] = make_double4(in->bias_vel[b][0], in->bias_vel[b][1], in->bias_angvel[b], 0.0);
  }
  for (int j = 0; j < n_joints; j++) st.jacc[j] = make_double2(in->joint_acc[j][0], in->joint_acc[j][1]);
  st.stamp = in->stamp;
  st.episode_steps = in->episode_steps;
  st.overflow = in->overflow;
  st.n_cache = in->n_cache;
  int last = 0;
  for (int k = 0; k < in->n_cache; k++) {
    CEntry e;
    memset(&e, 0, sizeof(e));
    e.a = (uint8_t)in->cache_shapes[k][0]; e.b = (uint8_t)in->cache_shapes[k][1];
    e.hash = in->cache_hash[k];
    e.stamp = in->stamp - in->cache_age[k];
    e.jn = in->cache_jn[k]; e.jt = in->cache_jt[k];
    st.cache[k] = e;
    if (in->cache_age[k] == 0) last++;
  }
  st.last_contacts = last;
  return nullptr;
}

#endif

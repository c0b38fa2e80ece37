// Host build of the PRODUCT's thread-per-environment physics step
// (magical_b200/csrc/mg_physics_tpe.h) so the CPU test-suite can run the very source the sm_100a
// kernel executes against the oracle, bit for bit, without a GPU.  Test infrastructure only.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../magical_b200/csrc/mg_physics_tpe.h"
#include "../../magical_b200/csrc/mg_state_io.h"

struct TpeHostEnv {
  DeviceScene ds;
  EnvState st;
  TpeLayout L;
  std::vector<double> words;
  std::vector<double> spill;
  std::vector<uint32_t> scratch;
  int use_spill;
};

static void host_reset(TpeHostEnv* e) {
  // state half of BaseEnv.reset, as reset_state()/k_reset in mg_finish.cu
  EnvState& st = e->st;
  const mg_scene_t& sc = e->ds.s;
  memset(&st, 0, sizeof(st));
  st.fresh = 1;
  for (int b = 0; b < MG_MAX_BODIES; b++) {
    double x = 0.0, y = 0.0, a = 0.0, cs = 1.0, sn = 0.0;
    if (b < sc.n_bodies) {
      x = sc.bodies[b].p0[0]; y = sc.bodies[b].p0[1]; a = sc.bodies[b].a0;
      mg_det_sincos(a, &sn, &cs);
    }
    st.P[b] = make_double4(x, y, a, 0.0);
    st.R[b] = make_double2(cs, sn);
  }
  for (int k = 0; k < MG_NCACHE; k++) st.cache[k].stamp = -100;
}

extern "C" {

TpeHostEnv* tpeh_create(const mg_scene_t* scene, int kcon, int use_spill, int nitems, int scratch_global) {
  TpeHostEnv* e = new TpeHostEnv();
  e->ds.s = *scene;
  if (mg_build_scene_aux(scene, &e->ds.aux) || !e->ds.aux.tpe_ok) { delete e; return nullptr; }
  e->L = tpe_make_layout(e->ds.aux.tpe_nslots, e->ds.aux.tpe_nblocks, scene->n_cgroups, scene->n_bpairs, kcon, nitems, scratch_global);
  e->scratch.assign((size_t)e->L.scratch_u32, 0u);
  e->words.assign((size_t)e->L.words, 0.0);
  e->spill.assign((size_t)TPE_MAX_CONTACTS * TPE_CON_WORDS, 0.0);
  e->use_spill = use_spill;
  host_reset(e);
  return e;
}
void tpeh_destroy(TpeHostEnv* e) { delete e; }
void tpeh_reset(TpeHostEnv* e) { host_reset(e); }
int tpeh_kcon(const TpeHostEnv* e) { return e->L.kcon; }
int tpeh_words(const TpeHostEnv* e) { return e->L.words; }

void tpeh_step(TpeHostEnv* e, int action) {
  Tpe<1> T;
  T.wd = e->words.data();
  T.wf = reinterpret_cast<float*>(e->words.data());
  T.wh = reinterpret_cast<uint16_t*>(e->words.data());
  T.L = e->L;
  T.spill = e->use_spill ? e->spill.data() : nullptr;
  T.slotmap = 0;
  T.static_slot = 0;
  T.bind_scratch(e->scratch.data());
  tpe_env_step<1>(T, &e->st, &e->ds, action, true);
  e->st.episode_steps++;
}

void tpeh_set_pose(TpeHostEnv* e, int body, double x, double y, double angle) {
  double sn, cs;
  mg_det_sincos(angle, &sn, &cs);
  e->st.P[body] = make_double4(x, y, angle, 0.0);
  e->st.R[body] = make_double2(cs, sn);
}

// same conversion as mg_get_state (mg_api.cu)
void tpeh_get_state(const TpeHostEnv* e, mg_state_t* out) { mg_state_export(e->st, e->ds.s, out); }
int tpeh_set_state(TpeHostEnv* e, const mg_state_t* in) {
  return mg_state_import(e->st, e->ds.s.n_bodies, e->ds.s.n_joints, e->ds.s.n_shapes, in) ? -1 : 0;
}

}  // extern "C"

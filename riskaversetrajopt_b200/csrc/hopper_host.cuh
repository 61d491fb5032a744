// placeholder
extern "C" {
int saa_set_samples_hopper(saa_handle *h, int32_t, double, const double *, const double *, const double *, void *) {
  return fail(h, SAA_ERR_ARG, "hopper: not built yet");
}
int saa_hopper_friction(saa_handle *h, int32_t, const double *, void *, void *, const double *, double *, void *) {
  return fail(h, SAA_ERR_ARG, "hopper: not built yet");
}
}

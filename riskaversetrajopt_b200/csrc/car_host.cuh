// placeholder
namespace {
template <typename T>
int launch_car_assemble(saa_handle *h, const double *, int, void *, void *, void *, int *, cudaStream_t) {
  return fail(h, SAA_ERR_ARG, "car: not built yet");
}
template <typename T>
int launch_car_rollout(saa_handle *h, const double *, void *, void *, double, double, double, double *, cudaStream_t) {
  return fail(h, SAA_ERR_ARG, "car: not built yet");
}
int car_write_constants_relaxed(saa_handle *h, int, void *, void *, void *, cudaStream_t) {
  return fail(h, SAA_ERR_ARG, "car: not built yet");
}
}
extern "C" {
int saa_set_params_car(saa_handle *h, const saa_car_params *p) {
  if (!h || !p) return fail(h, SAA_ERR_ARG, "NULL argument");
  h->cp = *p; h->params_set = true; return SAA_OK;
}
int saa_set_samples_car(saa_handle *h, const double *, const double *, const double *, const double *, void *) {
  return fail(h, SAA_ERR_ARG, "car: not built yet");
}
}

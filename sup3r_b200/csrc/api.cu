// libsup3r_b200 runtime glue: error convention, device init, descriptor validation.
#include <cstring>

#include "common.cuh"

namespace s3 {

static thread_local char g_err[512] = "";
static int g_sm_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return S3_ERR_CUDA;
}

int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_sm_count = n;
    else
      return 148;
  }
  return g_sm_count;
}

int make_geom(const s3_conv_desc* d, ConvGeom* g) {
  S3_REQUIRE(d != nullptr, "conv descriptor is null");
  S3_REQUIRE(d->ndim == 2 || d->ndim == 3, "conv ndim must be 2 or 3, got %d", d->ndim);
  S3_REQUIRE(d->n > 0 && d->cin > 0 && d->cout > 0, "conv n/cin/cout must be positive");
  memset(g, 0, sizeof(*g));
  g->ndim = d->ndim;
  g->n = d->n;
  g->cin = d->cin;
  g->cout = d->cout;
  g->pad_mode = d->pad_mode;
  g->act = d->act;
  g->alpha = d->alpha;
  S3_REQUIRE(d->pad_mode >= 0 && d->pad_mode <= 2, "bad pad_mode %d", d->pad_mode);
  for (int i = 0; i < 3; ++i) {
    g->in[i] = d->in_dims[i];
    g->k[i] = d->ksize[i];
    g->st[i] = d->stride[i];
    g->pl[i] = d->pad_lo[i];
    g->ph[i] = d->pad_hi[i];
    g->rep[i] = d->out_repeat[i] < 1 ? 1 : d->out_repeat[i];
    S3_REQUIRE(g->in[i] > 0 && g->k[i] > 0 && g->st[i] > 0 && g->pl[i] >= 0 && g->ph[i] >= 0,
               "bad conv geometry on dim %d", i);
    if (d->pad_mode == S3_PAD_REFLECT)
      S3_REQUIRE(g->pl[i] < g->in[i] && g->ph[i] < g->in[i],
                 "REFLECT padding %d/%d needs extent > pad (dim %d extent %d)", g->pl[i], g->ph[i],
                 i, g->in[i]);
    if (d->pad_mode == S3_PAD_SYMMETRIC)
      S3_REQUIRE(g->pl[i] <= g->in[i] && g->ph[i] <= g->in[i], "SYMMETRIC padding too large");
    const int span = g->in[i] + g->pl[i] + g->ph[i] - g->k[i];
    S3_REQUIRE(span >= 0, "conv input extent %d (+pad %d,%d) smaller than kernel %d on dim %d",
               g->in[i], g->pl[i], g->ph[i], g->k[i], i);
    g->od[i] = span / g->st[i] + 1;
  }
  if (d->ndim == 2)
    S3_REQUIRE(g->in[0] == 1 && g->k[0] == 1 && g->st[0] == 1 && g->pl[0] == 0 && g->ph[0] == 0,
               "2-D conv must have z extent / kernel / stride 1 and no z padding");
  g->r = d->d2s < 1 ? 1 : d->d2s;
  g->m = d->d2t < 1 ? 1 : d->d2t;
  g->roll = d->t_roll;
  S3_REQUIRE(g->m == 1 || d->ndim == 3, "depth_to_time needs a 3-D conv");
  g->res_pre = d->res_pre_act ? 1 : 0;
  g->ctotal = d->cout_total > 0 ? d->cout_total : g->cout;
  g->cbase = d->cout_base;
  S3_REQUIRE(g->cbase >= 0 && g->cbase + g->cout <= g->ctotal,
             "channel slice [%d, %d) exceeds cout_total %d", g->cbase, g->cbase + g->cout, g->ctotal);
  S3_REQUIRE(g->ctotal % (g->r * g->r * g->m) == 0,
             "cout %d not divisible by d2s^2 * d2t = %d", g->ctotal, g->r * g->r * g->m);
  g->cmap = g->ctotal / (g->r * g->r * g->m);
  if (d->ndim == 3) {
    g->fd[0] = g->od[0] * g->r * g->rep[0];
    g->fd[1] = g->od[1] * g->r * g->rep[1];
    g->fd[2] = g->od[2] * g->m * g->rep[2];
  } else {
    g->fd[0] = 1;
    g->fd[1] = g->od[1] * g->r * g->rep[1];
    g->fd[2] = g->od[2] * g->r * g->rep[2];
    S3_REQUIRE(g->rep[0] == 1, "2-D conv cannot repeat z");
  }
  g->cstride = d->out_cstride > 0 ? d->out_cstride : g->cmap;
  g->coff = d->out_coffset;
  S3_REQUIRE(g->coff >= 0 && g->coff + g->cmap <= g->cstride,
             "out_coffset %d + channels %d exceeds out_cstride %d", g->coff, g->cmap, g->cstride);
  return S3_OK;
}

}  // namespace s3

using namespace s3;

extern "C" const char* s3_last_error(void) { return g_err; }
extern "C" int s3_version(void) { return 100; }

extern "C" int s3_init(int device) {
  S3_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  S3_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("sup3r_b200 needs an sm_100 (B200) device; device %d is sm_%d%d", device, prop.major,
              prop.minor);
    return S3_ERR_UNSUPPORTED;
  }
  g_sm_count = prop.multiProcessorCount;
  return S3_OK;
}

extern "C" int s3_sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  return n;
}

extern "C" int s3_conv_out_dims(const s3_conv_desc* d, int32_t conv_dims[3], int32_t out_dims[3],
                                int32_t* out_channels) {
  ConvGeom g;
  int rc = make_geom(d, &g);
  if (rc) return rc;
  for (int i = 0; i < 3; ++i) {
    if (conv_dims) conv_dims[i] = g.od[i];
    if (out_dims) out_dims[i] = g.fd[i];
  }
  if (out_channels) *out_channels = g.cmap;
  return S3_OK;
}

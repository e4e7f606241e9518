// api.cu -- handle lifetime of libmkb200.so (see include/mkb200.h).
#include "common.cuh"

extern "C" {

int mkb_version(void) { return MKB_VERSION; }

int mkb_create(int device, mkb_handle_t *out) {
    if (!out) return MKB_ERR_BAD_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return MKB_ERR_CUDA;  // no CPU fallback: the library is useless without a CUDA device
    }
    if (device < 0 || device >= n) return MKB_ERR_BAD_ARG;
    mkb_ctx *h = new (std::nothrow) mkb_ctx();
    if (!h) return MKB_ERR_NOMEM;
    h->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    *out = h;
    return MKB_OK;
}

int mkb_destroy(mkb_handle_t h) {
    if (!h) return MKB_ERR_BAD_ARG;
    {
        mkb::DeviceGuard g(h->device);
        for (auto &s : h->scratch)
            if (s.ptr) cudaFree(s.ptr);
        for (auto &e : h->ev)
            if (e) cudaEventDestroy(e);
        if (h->order_ev) cudaEventDestroy(h->order_ev);
        for (auto &e : h->stage_ev)
            if (e) cudaEventDestroy(e);
        for (auto &b : h->host_stage)
            if (b) cudaFreeHost(b);
        for (auto &e : h->aux_ev)
            if (e) cudaEventDestroy(e);
        if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
        if (h->aux_stream2) cudaStreamDestroy(h->aux_stream2);
    }
    delete h;
    return MKB_OK;
}

const char *mkb_last_error(mkb_handle_t h) { return h ? h->err.c_str() : "null handle"; }

int64_t mkb_launch_count(mkb_handle_t h) { return h ? h->launches : -1; }

const char *mkb_last_kernel(mkb_handle_t h) { return h ? h->last_kernel : ""; }

int mkb_set_timing(mkb_handle_t h, int on) {
    MKB_ENTER(h);
    if (on && !h->ev[0])
        for (auto &e : h->ev) MKB_CUDA(h, cudaEventCreate(&e));
    h->timing = on != 0;
    return MKB_OK;
}

int mkb_get_timing(mkb_handle_t h, float *prep_ms, float *main_ms) {
    MKB_ENTER(h);
    if (!h->ev[0]) return mkb::fail(h, MKB_ERR_BAD_ARG, "timing was never enabled");
    MKB_CUDA(h, cudaEventSynchronize(h->ev[2]));
    if (prep_ms) MKB_CUDA(h, cudaEventElapsedTime(prep_ms, h->ev[0], h->ev[1]));
    if (main_ms) MKB_CUDA(h, cudaEventElapsedTime(main_ms, h->ev[1], h->ev[2]));
    return MKB_OK;
}

}  // extern "C"

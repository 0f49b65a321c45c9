// Host-buffer entry points of libctc_b200.so (include/ctc_b200.h): the call a host without device tensors makes.
// The batch is cut into slices and three streams form a pipeline: all host->device copies go back to back on the first
// (the PCIe link never idles between slices), each slice's kernels run on the second as soon as that slice's event
// fires, and the loss (and, when asked for, the gradient) of a finished slice travels device->host on the third, so the
// two PCIe directions run concurrently and everything but the last slice's kernel and read-back hides under the copies.
#include <initializer_list>
#include <new>

#include "common.cuh"

struct ctcb200_host_ctx {
  ctcb200_desc desc;
  int device;
  int num_slices;
  cudaStream_t streams[3];      // [0] copies host -> device, [1] kernels, [2] copies device -> host
  cudaEvent_t* landed;          // [num_slices] slice i is on the device
  cudaEvent_t* done;            // [num_slices] slice i's kernels have finished
  float* d_logits;
  float* d_grad;
  int32_t* d_labels;
  int32_t* d_label_length;
  int32_t* d_logit_length;
  float* d_loss;
  void* ws[2];
  size_t ws_bytes;
};

extern "C" {

void ctcb200_host_destroy(ctcb200_host_ctx* c) {
  if (c == nullptr) return;
  cudaSetDevice(c->device);
  for (int i = 0; i < 3; ++i)
    if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
  for (int i = 0; i < 2; ++i)
    if (c->ws[i]) cudaFree(c->ws[i]);
  for (cudaEvent_t* ev : {c->landed, c->done})
    if (ev) {
      for (int i = 0; i < c->num_slices; ++i)
        if (ev[i]) cudaEventDestroy(ev[i]);
      delete[] ev;
    }
  cudaFree(c->d_logits); cudaFree(c->d_grad); cudaFree(c->d_labels); cudaFree(c->d_label_length);
  cudaFree(c->d_logit_length); cudaFree(c->d_loss);
  delete c;
}

int ctcb200_host_create(const ctcb200_desc* desc, int device, int num_slices, ctcb200_host_ctx** out) {
  if (desc == nullptr || out == nullptr) return CTCB200_ERR_NULL_POINTER;
  *out = nullptr;
  if (desc->flags & CTCB200_TIME_MAJOR) return CTCB200_ERR_BAD_DESCRIPTOR;   // batch slices must be contiguous
  if (num_slices < 1) num_slices = 1;
  if (desc->B > 0 && num_slices > desc->B) num_slices = desc->B;
  const int slice_b = desc->B > 0 ? (desc->B + num_slices - 1) / num_slices : 0;
  // The training-call workspace is NOT monotonic in the batch size (a short tail slice of a narrow vocabulary falls back
  // from the fused kernel to the larger staged scratch, see ctc_b200.h): size it for the full slice and for the tail.
  ctcb200_desc sd = *desc;
  sd.B = slice_b;
  size_t ws_bytes = ctcb200_workspace_bytes(&sd, CTCB200_WS_LOSS_GRAD_LOGITS);
  if (ws_bytes == 0 && slice_b > 0) return CTCB200_ERR_BAD_DESCRIPTOR;
  if (slice_b > 0 && desc->B % slice_b != 0) {
    sd.B = desc->B % slice_b;
    const size_t tail = ctcb200_workspace_bytes(&sd, CTCB200_WS_LOSS_GRAD_LOGITS);
    if (tail > ws_bytes) ws_bytes = tail;
  }
  ctcb200_host_ctx* c = new (std::nothrow) ctcb200_host_ctx();
  if (c == nullptr) return CTCB200_ERR_CUDA;
  c->desc = *desc; c->device = device; c->num_slices = num_slices; c->ws_bytes = ws_bytes;
  const size_t n = (size_t)desc->B * desc->T * desc->V;
  bool ok = cudaSetDevice(device) == cudaSuccess;
  for (int i = 0; i < 3 && ok; ++i) ok = ok && cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc(&c->ws[0], ws_bytes ? ws_bytes : 256) == cudaSuccess;     // kernels run in order: one workspace
  c->landed = new (std::nothrow) cudaEvent_t[num_slices]();
  c->done = new (std::nothrow) cudaEvent_t[num_slices]();
  ok = ok && c->landed != nullptr && c->done != nullptr;
  for (int i = 0; i < num_slices && ok; ++i)
    ok = ok && cudaEventCreateWithFlags(&c->landed[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_logits, n ? n * 4 : 256) == cudaSuccess;     // (sized for fp32; bf16 rows use half of it)
  ok = ok && cudaMalloc(&c->d_grad, n ? n * 4 : 256) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_labels, (size_t)desc->B * desc->Lw * 4 + 256) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_label_length, (size_t)desc->B * 4 + 256) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_logit_length, (size_t)desc->B * 4 + 256) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_loss, (size_t)desc->B * 4 + 256) == cudaSuccess;
  if (!ok) {
    (void)cudaGetLastError();
    ctcb200_host_destroy(c);
    return CTCB200_ERR_CUDA;
  }
  *out = c;
  return CTCB200_OK;
}

float* ctcb200_host_grad_device_ptr(ctcb200_host_ctx* c) { return c ? c->d_grad : nullptr; }

int ctcb200_host_loss_grad(ctcb200_host_ctx* c, const float* host_logits, const int32_t* host_labels,
                           const int32_t* host_label_length, const int32_t* host_logit_length, float* host_loss,
                           float* host_grad_logits) {
  if (c == nullptr) return CTCB200_ERR_NULL_POINTER;
  const ctcb200_desc& d = c->desc;
  if (d.B == 0) return CTCB200_OK;
  if (host_label_length == nullptr || host_logit_length == nullptr || host_loss == nullptr ||
      (d.T > 0 && host_logits == nullptr) || (d.Lw > 0 && host_labels == nullptr))
    return CTCB200_ERR_NULL_POINTER;
  if (cudaSetDevice(c->device) != cudaSuccess) return CTCB200_ERR_CUDA;
  const int slice_b = (d.B + c->num_slices - 1) / c->num_slices;
  const size_t tv = (size_t)d.T * d.V;
  // element sizes of the logits / gradient rows: host and device buffers hold the same format (CTCB200_LOGITS_BF16 /
  // CTCB200_GRAD_BF16), so bf16 logits halve the bytes that cross PCIe
  const size_t es_in = (d.flags & CTCB200_LOGITS_BF16) ? 2 : 4, es_out = (d.flags & CTCB200_GRAD_BF16) ? 2 : 4;
  const char* h_in = reinterpret_cast<const char*>(host_logits);
  char* h_out = reinterpret_cast<char*>(host_grad_logits);
  char *dv_in = reinterpret_cast<char*>(c->d_logits), *dv_out = reinterpret_cast<char*>(c->d_grad);
  int rc = CTCB200_OK;
  cudaStream_t copy = c->streams[0], run = c->streams[1], back = c->streams[2];
  for (int i = 0, b0 = 0; b0 < d.B; ++i, b0 += slice_b) {
    const int nb = (d.B - b0 < slice_b) ? d.B - b0 : slice_b;
    ctcb200_desc sd = d;
    sd.B = nb;
    bool ok = true;
    if (tv) ok = ok && cudaMemcpyAsync(dv_in + b0 * tv * es_in, h_in + b0 * tv * es_in, nb * tv * es_in, cudaMemcpyHostToDevice, copy) == cudaSuccess;
    if (d.Lw) ok = ok && cudaMemcpyAsync(c->d_labels + (size_t)b0 * d.Lw, host_labels + (size_t)b0 * d.Lw, (size_t)nb * d.Lw * 4, cudaMemcpyHostToDevice, copy) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(c->d_label_length + b0, host_label_length + b0, (size_t)nb * 4, cudaMemcpyHostToDevice, copy) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(c->d_logit_length + b0, host_logit_length + b0, (size_t)nb * 4, cudaMemcpyHostToDevice, copy) == cudaSuccess;
    ok = ok && cudaEventRecord(c->landed[i], copy) == cudaSuccess;
    ok = ok && cudaStreamWaitEvent(run, c->landed[i], 0) == cudaSuccess;
    if (!ok) { rc = CTCB200_ERR_CUDA; break; }
    rc = ctcb200_loss_grad(&sd, reinterpret_cast<const float*>(dv_in + b0 * tv * es_in), c->d_labels + (size_t)b0 * d.Lw,
                           c->d_label_length + b0, c->d_logit_length + b0, nullptr, c->d_loss + b0,
                           reinterpret_cast<float*>(dv_out + b0 * tv * es_out), nullptr,
                           c->ws[0], c->ws_bytes, run);
    if (rc != CTCB200_OK) break;
    ok = cudaEventRecord(c->done[i], run) == cudaSuccess && cudaStreamWaitEvent(back, c->done[i], 0) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(host_loss + b0, c->d_loss + b0, (size_t)nb * 4, cudaMemcpyDeviceToHost, back) == cudaSuccess;
    if (host_grad_logits && tv)
      ok = ok && cudaMemcpyAsync(h_out + b0 * tv * es_out, dv_out + b0 * tv * es_out, nb * tv * es_out, cudaMemcpyDeviceToHost, back) == cudaSuccess;
    if (!ok) { rc = CTCB200_ERR_CUDA; break; }
  }
  for (int i = 0; i < 3; ++i)
    if (cudaStreamSynchronize(c->streams[i]) != cudaSuccess) rc = CTCB200_ERR_CUDA;
  if (rc == CTCB200_ERR_CUDA) (void)cudaGetLastError();
  return rc;
}

}  // extern "C"

// Decoder entry points (placeholder until decoder.cu lands; every call fails loudly).
#include "../../include/lrpcap.h"
#include "common.cuh"
using namespace lrpcap;
extern "C" {
#define NOT_YET(name) set_last_error(name ": decoder not built yet"); return kErrUnsupported
int lrpcap_decoder_create(lrpcap_decoder_t**, const lrpcap_decoder_weights_t*, int, int) { NOT_YET("decoder_create"); }
int lrpcap_decoder_destroy(lrpcap_decoder_t*) { return kOk; }
int lrpcap_decoder_forward(lrpcap_decoder_t*, const float*, int, int, int*, int, int, int, void*) { NOT_YET("decoder_forward"); }
int lrpcap_decoder_relevance(lrpcap_decoder_t*, const int*, const int*, int, float*, double*, float*, void*) { NOT_YET("decoder_relevance"); }
int lrpcap_decoder_backward(lrpcap_decoder_t*, const int*, const int*, int, float*, double*, void*) { NOT_YET("decoder_backward"); }
int lrpcap_decoder_caption_logits(lrpcap_decoder_t*, double*) { NOT_YET("decoder_caption_logits"); }
long long lrpcap_decoder_launches(lrpcap_decoder_t*) { return 0; }
int lrpcap_explain_batch_host(lrpcap_encoder_t*, lrpcap_decoder_t*, const float*, int, int*, int, int, int, int, int, float, float, float, int, float*, void*) { NOT_YET("explain_batch_host"); }
}

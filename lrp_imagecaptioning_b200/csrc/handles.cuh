// Opaque handle layouts behind the C ABI.
#pragma once
#include "decoder.cuh"
#include "encoder.cuh"
struct lrpcap_encoder {
  lrpcap::Encoder* impl;
};
struct lrpcap_decoder {
  lrpcap::Decoder* impl;
};

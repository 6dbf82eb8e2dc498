// Opaque handle layouts behind the C ABI.
#pragma once
#include "decoder.cuh"
#include "encoder.cuh"
struct lrpcap_encoder {
  lrpcap::Encoder* impl;
  // device staging of lrpcap_explain_batch_host (images, head relevance, pixel maps); owned by the handle so that
  // concurrent calls on different handles / host threads do not share or re-allocate it
  lrpcap::DevBuf stage_img, stage_head, stage_pix;
  ~lrpcap_encoder() {
    stage_img.release();
    stage_head.release();
    stage_pix.release();
  }
};
struct lrpcap_decoder {
  lrpcap::Decoder* impl;
};

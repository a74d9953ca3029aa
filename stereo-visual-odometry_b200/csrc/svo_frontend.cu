// TEMPORARY stubs (replaced by the real implementation in the next milestone)
#include "svo_common.cuh"
extern "C" int svo_fast_detect(svo_ctx* c, const uint8_t*, int, int, size_t, int, int, svo_keypoint*, int, int*) { return c ? c->fail(SVO_ERR_INVALID, "not implemented") : SVO_ERR_INVALID; }
extern "C" int svo_lk_track(svo_ctx* c, const uint8_t*, const uint8_t*, int, int, size_t, const float*, int, float*, uint8_t*, float*, int, int, int, double, double) { return c ? c->fail(SVO_ERR_INVALID, "not implemented") : SVO_ERR_INVALID; }
extern "C" int svo_frontend_configure(svo_ctx* c, int, int, int, int) { return c ? c->fail(SVO_ERR_INVALID, "not implemented") : SVO_ERR_INVALID; }
extern "C" int svo_frontend_reset(svo_ctx* c) { return c ? c->fail(SVO_ERR_INVALID, "not implemented") : SVO_ERR_INVALID; }
extern "C" int svo_frontend_upload(svo_ctx* c, const uint8_t*, const uint8_t*, int, size_t) { return c ? c->fail(SVO_ERR_INVALID, "not implemented") : SVO_ERR_INVALID; }
extern "C" int svo_frontend_run(svo_ctx* c) { return c ? c->fail(SVO_ERR_INVALID, "not implemented") : SVO_ERR_INVALID; }
extern "C" int svo_frontend_download(svo_ctx* c, svo_frontend_out*) { return c ? c->fail(SVO_ERR_INVALID, "not implemented") : SVO_ERR_INVALID; }
extern "C" int svo_frontend_batch(svo_ctx* c, const uint8_t*, const uint8_t*, int, size_t, svo_frontend_out*) { return c ? c->fail(SVO_ERR_INVALID, "not implemented") : SVO_ERR_INVALID; }

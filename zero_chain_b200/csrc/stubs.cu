// TEMPORARY: entry points not implemented yet fail loudly (replaced as ntt.cu / groth16.cu land).
#include "internal.h"
#define NOT_YET(name) { zk_set_error(name ": not implemented yet"); return ZK_ERR_INVALID; }
extern "C" int zk_ntt_fr(zk_ctx *, uint64_t *, unsigned, int) NOT_YET("zk_ntt_fr")
extern "C" int zk_ntt_fr_device(zk_ctx *, void *, unsigned, int) NOT_YET("zk_ntt_fr_device")
extern "C" int zk_params_load(zk_ctx *, const uint8_t *, size_t, int, zk_params **) NOT_YET("zk_params_load")
extern "C" void zk_params_free(zk_params *) {}
extern "C" int zk_params_counts(const zk_params *, uint64_t *) NOT_YET("zk_params_counts")
extern "C" int zk_groth16_prove(zk_ctx *, const zk_params *, const uint64_t *, const uint64_t *, const uint64_t *, size_t, const uint64_t *, size_t,
                                const uint64_t *, size_t, const uint8_t *, const uint8_t *, const uint8_t *, const uint64_t *, const uint64_t *, uint8_t *) NOT_YET("zk_groth16_prove")
extern "C" int zk_groth16_prove_batch(zk_ctx *, const zk_params *, size_t, const uint64_t *, const uint64_t *, const uint64_t *, size_t, const uint64_t *, size_t,
                                      const uint64_t *, size_t, const uint8_t *, const uint8_t *, const uint8_t *, const uint64_t *, const uint64_t *, uint8_t *) NOT_YET("zk_groth16_prove_batch")

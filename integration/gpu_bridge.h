/* gpu_bridge.h — the one call the two edited objective callbacks of the reference make (see gpu_bridge.cpp). */
#ifndef CAFE_GPU_BRIDGE_H
#define CAFE_GPU_BRIDGE_H
extern "C" {
#include "family.h"   /* pCafeParam (libtree/family.h:104-172) */
}
/* One objective evaluation on the GPU: what
 *     reset_birthdeath_cache(...); score = get_posterior(param->pfamily, param->pcafe, pr); cafe_free_birthdeath_cache(...)
 * computes in __cafe_best_lambda_search (cafe/lambda.cpp:744-761) and cafe_best_lambda_mu_search (cafe/lambdamu.cpp:341-356).
 * Throws std::runtime_error with the reference's message when a family has likelihood 0 (cafe/lambda.cpp:715-720). */
double gpu_bridge_score(pCafeParam param);
#endif

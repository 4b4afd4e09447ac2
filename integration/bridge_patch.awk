# bridge_patch.awk — the two edits INTEGRATION.md §2 asks a maintainer of the reference to make, applied by oracle/Makefile
# (target `bridge`) to copies of cafe/lambda.cpp and cafe/lambdamu.cpp under oracle/_ref/bridge/ (build outputs, git-ignored):
# inside __cafe_best_lambda_search (cafe/lambda.cpp:726-769) and cafe_best_lambda_mu_search (cafe/lambdamu.cpp:323-367) the
# statements  reset_birthdeath_cache(...); score = get_posterior(...); cafe_free_birthdeath_cache(pcafe);  become
# score = gpu_bridge_score(param);  Everything else of the reference is compiled unmodified.
NR == 1 { print "#include \"gpu_bridge.h\"" }
/^double __cafe_best_lambda_search\(/ || /^double cafe_best_lambda_mu_search\(/ { infn = 1 }
infn && /^}/ { infn = 0 }
infn && /reset_birthdeath_cache\(param->pcafe, param->parameterized_k_value, &param->family_size\);/ { edits++; next }
infn && /cafe_free_birthdeath_cache\(pcafe\);/ { edits++; next }
infn && /score = get_posterior\(param->pfamily, param->pcafe, pr\);/ { sub(/get_posterior\(param->pfamily, param->pcafe, pr\)/, "gpu_bridge_score(param)"); edits++ }
{ print }
END { if (edits != 3) { print "bridge_patch.awk: expected 3 edits, made " edits > "/dev/stderr"; exit 1 } }

/* cnb_result.h -- host-side result objects (the catNetwork / catNetworkEvaluate slots the R shim exports). */
#ifndef CNB_RESULT_H
#define CNB_RESULT_H
#include <vector>
#include "cnb_internal.h"

/* one scored family = one node's conditional table inside a returned network (PROB_LIST, problist.h:37-46) */
struct CnbFamily {
	int32_t child = 0;                 /* order position */
	int32_t d = 0;
	int32_t pars[CNB_MAXP] = {0};      /* order positions, table order */
	int32_t cc = 0;                    /* categories of the child */
	int32_t pc[CNB_MAXP] = {0};        /* categories of the parents */
	std::vector<int32_t> counts;       /* reference table layout */
	double loglik = 0, prior = 0;      /* PROB_LIST::loglik (sample-averaged), ::priorlik */
	int32_t ncount = 0;                /* PROB_LIST::sampleSize */
};

struct CnbNet {
	int32_t complexity = 0;
	double loglik = 0;
	int32_t order_id = 0;              /* which order its positions refer to */
	std::vector<int32_t> fam;          /* [N] index into cnb_result::fams */
};

struct cnb_result {
	int32_t N = 0;
	std::vector<std::vector<int32_t>> orders;   /* 1-based labels per position */
	std::vector<std::vector<int32_t>> order_cats; /* categories per position, one entry per order */
	std::vector<CnbFamily> fams;
	std::vector<CnbNet> nets;                   /* ascending complexity */
	/* SA bookkeeping (rcatnet_sa.cpp) */
	std::vector<int32_t> sa_order;
	int32_t niter = 0, naccept = 0;
	std::vector<double> trace;
	std::vector<int32_t> trace_accept;
};

#endif

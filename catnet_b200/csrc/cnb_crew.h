/* cnb_crew.h -- a small crew of host threads: one per device of a cnb_mctx (cnb_multi.cu), one per concurrently evaluated
 * order of cnSearchSA / cnSearchHist (cnb_api.cu).  What the reference's worker pool is (src/thread.h, src/pthread.cpp,
 * rcatnet_sa.cpp:475-652): a fixed set of threads that each take one job of a batch, the caller waiting for all of them. */
#ifndef CNB_CREW_H
#define CNB_CREW_H
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "cnb_internal.h"

/* Worker g (1 <= g < G) sleeps on a condition variable; run(n, f) wakes workers 1..n-1, runs f(0) on the calling
 * thread (R's thread: R API callbacks such as the uniform generator stay there) and returns when all are done. */
struct Crew {
	int G = 0;
	std::vector<std::thread> th;
	std::mutex mu;
	std::condition_variable cv_go, cv_done;
	std::function<int(int)> job;
	unsigned long long epoch = 0;
	int active = 0, pending = 0;
	bool quit = false;
	std::vector<int> rc;
	std::vector<std::string> err;

	void start(int g_count) {
		G = g_count;
		rc.assign(G, CNB_OK);
		err.assign(G, std::string());
		for (int g = 1; g < G; g++) th.emplace_back([this, g]() { loop(g); });
	}
	void loop(int g) {
		unsigned long long seen = 0;
		for (;;) {
			std::unique_lock<std::mutex> lk(mu);
			cv_go.wait(lk, [&] { return quit || epoch != seen; });
			if (quit) return;
			seen = epoch;
			if (g >= active) continue;
			std::function<int(int)> f = job;
			lk.unlock();
			int r = f(g);
			std::string e = r != CNB_OK ? cnb_last_error() : "";
			lk.lock();
			rc[g] = r;
			err[g] = e;
			if (--pending == 0) cv_done.notify_one();
		}
	}
	/* f(g) for g in [0, n); the first failure (lowest g) is reported with its message */
	int run(int n, const std::function<int(int)> &f) {
		n = std::max(1, std::min(n, G));
		if (n > 1) {
			std::lock_guard<std::mutex> lk(mu);
			job = f;
			active = n;
			pending = n - 1;
			epoch++;
		}
		if (n > 1) cv_go.notify_all();
		rc[0] = f(0);
		if (rc[0] != CNB_OK) err[0] = cnb_last_error();
		if (n > 1) {
			std::unique_lock<std::mutex> lk(mu);
			cv_done.wait(lk, [&] { return pending == 0; });
		}
		for (int g = 0; g < n; g++)
			if (rc[g] != CNB_OK) {
				cnb_set_error("%s%s", err[g].c_str(), g ? " (on a secondary device)" : "");
				return rc[g];
			}
		return CNB_OK;
	}
	void stop() {
		{
			std::lock_guard<std::mutex> lk(mu);
			quit = true;
		}
		cv_go.notify_all();
		for (auto &t : th) t.join();
		th.clear();
	}
};


#endif

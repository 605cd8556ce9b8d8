// RunMerger.h -- merge scheduling over device-resident runs.
//
// Replaces KMerFileMergeHandler (KMerFileMergeHandler.cpp:23-123), whose polling thread
// hangs when InputComplete() arrives late or when there is a single run.
// Same shape -- AddRun / InputComplete / result -- but synchronous and terminating for
// 0, 1 or N runs: whenever noOfMergersAtOnce runs are pending they are merged on the GPU
// (kc_merge_runs, the merge-path kernel), and InputComplete() merges whatever is left.
#pragma once

#include <stdint.h>

#include <vector>

#include "../../include/kc_api.h"

class RunMerger {
public:
    RunMerger(kc_ctx *ctx, uint32_t noOfMergersAtOnce) : _ctx(ctx), _fanIn(noOfMergersAtOnce < 2 ? 2 : noOfMergersAtOnce) {}
    ~RunMerger() { for (kc_run *r : _pending) kc_run_free(_ctx, r); }
    int AddRun(kc_run *run) {                       // takes ownership
        _pending.push_back(run);
        return _pending.size() >= _fanIn ? mergePending() : KC_OK;
    }
    // Merges everything that is left; *out is the final run (an empty run for no input).
    int InputComplete(kc_run **out) {
        int rc = mergePending();
        if (rc != KC_OK) return rc;
        if (_pending.empty()) return kc_merge_runs(_ctx, nullptr, 0, out);
        *out = _pending[0];
        _pending.clear();
        return KC_OK;
    }
    uint64_t merges() const { return _merges; }

private:
    int mergePending() {
        if (_pending.size() < 2) return KC_OK;
        kc_run *merged = nullptr;
        int rc = kc_merge_runs(_ctx, _pending.data(), (uint32_t)_pending.size(), &merged);
        if (rc != KC_OK) return rc;
        for (kc_run *r : _pending) kc_run_free(_ctx, r);
        _pending.assign(1, merged);
        _merges++;
        return KC_OK;
    }
    kc_ctx *_ctx;
    size_t _fanIn;
    std::vector<kc_run *> _pending;
    uint64_t _merges = 0;
};

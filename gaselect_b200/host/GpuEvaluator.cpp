// GpuEvaluator.cpp -- see GpuEvaluator.h.
#include "GpuEvaluator.h"

#include <cstdlib>
#include <cstring>
#include <string>

namespace {
std::vector<int> g_defaultDevices;
bool g_defaultDevicesSet = false;

// GASELECT_B200_DEVICES: "all", or ordinals separated by commas
std::vector<int> devicesFromEnvironment() {
	std::vector<int> out;
	const char* v = std::getenv("GASELECT_B200_DEVICES");
	if (!v || !*v) return out;
	if (std::strcmp(v, "all") == 0) {
		int32_t n = 0;
		if (gsb_device_count(&n) == GSB_OK) for (int i = 0; i < n && i < GSB_MAX_DEVICES; ++i) out.push_back(i);
		return out;
	}
	const char* p = v;
	while (*p) {
		char* end = nullptr;
		const long d = std::strtol(p, &end, 10);
		if (end == p) break;
		out.push_back(static_cast<int>(d));
		p = (*end == ',') ? end + 1 : end;
	}
	return out;
}
}

void GpuEvaluator::setDefaultDevices(const std::vector<int>& devices) {
	g_defaultDevices = devices;
	g_defaultDevicesSet = true;
}

void GpuEvaluator::check(gsb_engine* e, int rc) {
	if (rc == GSB_OK) return;
	char msg[1024];
	gsb_last_error(e, msg, sizeof(msg));
	if (rc == GSB_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
	throw std::runtime_error(msg);
}

GpuEvaluator* GpuEvaluator::make(const gsb_config& cfg, const arma::mat& X, const arma::vec& y,
                                 const std::vector<uint32_t>* seed, VerbosityLevel verbosity) {
	std::shared_ptr<Engine> eng(new Engine());
	gsb_config full = cfg;
	const std::vector<int> devices = g_defaultDevicesSet ? g_defaultDevices : devicesFromEnvironment();
	if (devices.size() > GSB_MAX_DEVICES) throw std::invalid_argument("too many devices");
	full.num_devices = static_cast<int32_t>(devices.size());
	for (size_t i = 0; i < devices.size(); ++i) full.devices[i] = devices[i];
	check(nullptr, gsb_create(&eng->handle, &full));
	// arma::mat is column-major, exactly what gsb_set_data takes; the engine keeps its own copy
	check(eng->handle, gsb_set_data(eng->handle, X.memptr(), y.memptr(), static_cast<int32_t>(X.n_rows), static_cast<int32_t>(X.n_cols)));
	if (seed) {
		if (seed->size() != GSB_SEED_WORDS) throw std::invalid_argument("the seed must have RNG::SEED_SIZE words");
		check(eng->handle, gsb_init_segmentation(eng->handle, seed->data()));
	}
	return new GpuEvaluator(eng, verbosity);
}

static gsb_config baseConfig(int evaluator, int device) {
	gsb_config cfg;
	std::memset(&cfg, 0, sizeof(cfg));
	cfg.struct_size = static_cast<int32_t>(sizeof(cfg));
	cfg.device = device;
	cfg.evaluator = evaluator;
	cfg.sdfact = 1.0;
	return cfg;
}

GpuEvaluator* GpuEvaluator::makePLS(const arma::mat& X, const arma::vec& y, uint16_t numReplications, uint16_t maxNComp,
                                    const std::vector<uint32_t>& seed, VerbosityLevel verbosity, uint16_t innerSegments,
                                    uint16_t outerSegments, double testSetSize, double sdfact, double complexityPenalty,
                                    int device) {
	gsb_config cfg = baseConfig(GSB_EVAL_PLS_CV, device);
	cfg.num_replications = numReplications;
	cfg.max_ncomp = maxNComp;
	cfg.inner_segments = innerSegments;
	cfg.outer_segments = outerSegments;
	cfg.test_set_size = testSetSize;
	cfg.sdfact = sdfact;
	cfg.complexity_penalty = complexityPenalty;
	return make(cfg, X, y, &seed, verbosity);
}

GpuEvaluator* GpuEvaluator::makeFit(const arma::mat& X, const arma::vec& y, uint16_t maxNComp, const std::vector<uint32_t>& seed,
                                    VerbosityLevel verbosity, uint16_t numSegments, int statistic, double sdfact, int device) {
	gsb_config cfg = baseConfig(GSB_EVAL_PLS_FIT, device);
	cfg.max_ncomp = maxNComp;
	cfg.inner_segments = numSegments;
	cfg.statistic = statistic;
	cfg.sdfact = sdfact;
	return make(cfg, X, y, &seed, verbosity);
}

GpuEvaluator* GpuEvaluator::makeLM(const arma::mat& X, const arma::vec& y, int statistic, VerbosityLevel verbosity, int device) {
	gsb_config cfg = baseConfig(GSB_EVAL_LM, device);
	cfg.statistic = statistic;
	return make(cfg, X, y, nullptr, verbosity);
}

void GpuEvaluator::evaluateBatch(const std::vector<arma::uvec>& subsets, std::vector<double>& fitness, std::vector<int>& status) {
	const size_t n = subsets.size();
	std::vector<uint32_t> offsets(n + 1, 0u), idx;
	for (size_t i = 0; i < n; ++i) {
		for (arma::uword j = 0; j < subsets[i].n_elem; ++j) idx.push_back(static_cast<uint32_t>(subsets[i][j]));
		offsets[i + 1] = static_cast<uint32_t>(idx.size());
	}
	if (idx.empty()) idx.push_back(0u);
	fitness.assign(n, 0.0);
	std::vector<int32_t> st(n, 0);
	{
		std::lock_guard<std::mutex> guard(engine->lock);
		check(engine->handle, gsb_evaluate_indices(engine->handle, idx.data(), offsets.data(), static_cast<int32_t>(n),
		                                           fitness.data(), st.data()));
		// GSB_STATUS_FLAGGED is not a failure (the fitness is valid; a near-tie decided optNComp): counted, reported as OK
		for (size_t i = 0; i < n; ++i) {
			if (st[i] == GSB_STATUS_FLAGGED) { st[i] = GSB_STATUS_OK; ++engine->flagged; }
		}
	}
	status.assign(st.begin(), st.end());
}

void GpuEvaluator::evaluateBitsets(const uint64_t* parts, int words, size_t count, std::vector<double>& fitness, std::vector<int>& status) {
	fitness.assign(count, 0.0);
	std::vector<int32_t> st(count, 0);
	if (count > 0) {
		std::lock_guard<std::mutex> guard(engine->lock);
		check(engine->handle, gsb_evaluate_bitsets(engine->handle, parts, words, static_cast<int32_t>(count), fitness.data(), st.data()));
		for (size_t i = 0; i < count; ++i) {
			if (st[i] == GSB_STATUS_FLAGGED) { st[i] = GSB_STATUS_OK; ++engine->flagged; }
		}
	}
	status.assign(st.begin(), st.end());
}

size_t GpuEvaluator::flaggedCount() const {
	std::lock_guard<std::mutex> guard(engine->lock);
	return engine->flagged;
}

double GpuEvaluator::evaluate(arma::uvec& columnSubset) {
	std::vector<arma::uvec> one(1, columnSubset);
	std::vector<double> fitness;
	std::vector<int> status;
	evaluateBatch(one, fitness, status);
	switch (status[0]) {
	case GSB_STATUS_OK: return fitness[0];
	case GSB_STATUS_EMPTY: throw Evaluator::EvaluatorException("Can not evaluate empty variable subset"); // src/PLSEvaluator.cpp:74
	case GSB_STATUS_UNDERFLOW: throw Evaluator::EvaluatorException("Can not perform PLS regression: Underflow error"); // :339
	default: throw Evaluator::EvaluatorException("The system could not be solved"); // src/LMEvaluator.cpp:61
	}
}

Evaluator* GpuEvaluator::clone() const {
	// the reference clones deep-copy X and y per thread (src/PLSSimpls.cpp:24-26); here the clones share the
	// device-resident engine and serialise on it (fitness is a pure function of the subset)
	return new GpuEvaluator(engine, verbosity);
}

std::vector<arma::uvec> GpuEvaluator::getSegmentation() const {
	int32_t nvec = 0;
	int64_t total = 0;
	check(engine->handle, gsb_get_segmentation(engine->handle, nullptr, nullptr, &nvec, &total));
	std::vector<uint32_t> rows(static_cast<size_t>(total) + 1), offsets(static_cast<size_t>(nvec) + 1);
	check(engine->handle, gsb_get_segmentation(engine->handle, rows.data(), offsets.data(), &nvec, &total));
	std::vector<arma::uvec> seg;
	seg.reserve(static_cast<size_t>(nvec));
	for (int32_t v = 0; v < nvec; ++v) {
		arma::uvec s(offsets[v + 1] - offsets[v]);
		for (uint32_t i = offsets[v]; i < offsets[v + 1]; ++i) s[i - offsets[v]] = rows[i];
		seg.push_back(s);
	}
	return seg;
}

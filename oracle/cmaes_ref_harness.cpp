// cmaes_ref_harness.cpp -- C entry points around the reference's OWN CMA-ES solver and controller
// (CovarianceMatrixAdaptationEvolution/CmaEsSolverTorch.cpp, Controller.cpp), compiled unchanged from where they lie
// under /root/reference against the libtorch that ships inside the pip torch wheel (oracle/Makefile, target
// _ref/libcmaes_ref.so).  TEST INFRASTRUCTURE ONLY: it pins oracle/cmaes_oracle.py (the Python restatement) and,
// through it, openkitchen_b200/cmaes.py.  Nothing here is product code and nothing is copied from the reference: the
// solver's private state is read through the usual test trick of re-declaring `private` for one include.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <random>
#include <vector>

#include <torch/torch.h>

#define private public
#include "CmaEsSolverTorch.h" // reference
#undef private
#include "Controller.h" // reference

namespace
{
void copy_out(const torch::Tensor &t, float *dst)
{
    if (!dst)
        return;
    const torch::Tensor c = t.detach().to(torch::kCPU, torch::kFloat32).contiguous();
    std::memcpy(dst, c.data_ptr<float>(), sizeof(float) * static_cast<size_t>(c.numel()));
}
} // namespace

extern "C"
{
// torch::randn in CmaEsSolver::sample draws from the default CPU generator
void cmr_seed(uint64_t seed)
{
    torch::manual_seed(seed);
}

void *cmr_create(int num_params, int population, float sigma)
{
    std::cout.setstate(std::ios_base::failbit); // the constructor prints every weight (CmaEsSolverTorch.cpp:26-29)
    auto *s = new CmaEsSolver(num_params, population, torch::kCPU, sigma);
    std::cout.clear();
    return s;
}

void cmr_destroy(void *h)
{
    delete static_cast<CmaEsSolver *>(h);
}

// CmaEsSolver::sample(): population x num_params, row i = candidate i
void cmr_sample(void *h, float *out)
{
    auto      *s   = static_cast<CmaEsSolver *>(h);
    const auto pop = s->sample();
    for (size_t i = 0; i < pop.size(); ++i)
        copy_out(pop[i], out + i * static_cast<size_t>(s->num_params_));
}

// CmaEsSolver::tell(): solutions as returned by sample (float32), one fitness each
void cmr_tell(void *h, const float *solutions, const float *fitness)
{
    auto                           *s = static_cast<CmaEsSolver *>(h);
    std::vector<SolutionAndFitness> v;
    v.reserve(static_cast<size_t>(s->population_size_));
    for (int i = 0; i < s->population_size_; ++i)
    {
        torch::Tensor x = torch::empty({s->num_params_}, torch::kFloat32);
        std::memcpy(x.data_ptr<float>(), solutions + static_cast<size_t>(i) * s->num_params_, sizeof(float) * s->num_params_);
        v.push_back({x, fitness[i]});
    }
    s->tell(v);
}

// every piece of solver state; any destination may be NULL.  scalars[8] = sigma, mu_eff, c_sigma, d_sigma, c_c, c_1,
// c_mu, chiN.  dtypes[5] = scalar type ids (c10::ScalarType) of mean, C, p_sigma, p_c, D as the reference left them
void cmr_get(void *h, float *mean, float *C, float *p_sigma, float *p_c, float *B, float *D, float *weights, float *scalars,
             int32_t *dtypes)
{
    auto *s = static_cast<CmaEsSolver *>(h);
    copy_out(s->param_mean_, mean);
    copy_out(s->C_, C);
    copy_out(s->p_sigma_, p_sigma);
    copy_out(s->p_c_, p_c);
    copy_out(s->B_, B);
    copy_out(s->D_, D);
    copy_out(s->weights_, weights);
    if (scalars)
    {
        const float v[8] = {s->sigma_, s->mu_eff_, s->c_sigma_, s->d_sigma_, s->c_c_, s->c_1_, s->c_mu_, s->chiN_};
        std::memcpy(scalars, v, sizeof v);
    }
    if (dtypes)
    {
        const torch::Tensor *t[5] = {&s->param_mean_, &s->C_, &s->p_sigma_, &s->p_c_, &s->D_};
        for (int i = 0; i < 5; ++i)
            dtypes[i] = static_cast<int32_t>(t[i]->scalar_type());
    }
}

void cmr_get_best(void *h, float *out)
{
    copy_out(static_cast<CmaEsSolver *>(h)->get_best_solution(), out);
}

// Controller (Controller.cpp:3-23) with the flat parameter vector of one candidate (set_params, Controller.cpp:45-54)
int cmr_controller_num_params(int inputs, int hidden, int outputs)
{
    Controller c(inputs, hidden, outputs);
    return static_cast<int>(c.count_params());
}

void cmr_controller_forward(int inputs, int hidden, int outputs, const float *flat, const float *x, int batch, float *out)
{
    torch::NoGradGuard no_grad;
    Controller         c(inputs, hidden, outputs);
    const int64_t      np = c.count_params();
    for (int b = 0; b < batch; ++b)
    { // one candidate at a time, like CmaEsAgent::updateAction (main_torch.cpp:61-71)
        torch::Tensor p = torch::empty({np}, torch::kFloat32);
        std::memcpy(p.data_ptr<float>(), flat + static_cast<size_t>(b) * np, sizeof(float) * np);
        c.set_params(p);
        torch::Tensor in = torch::empty({inputs}, torch::kFloat32);
        std::memcpy(in.data_ptr<float>(), x + static_cast<size_t>(b) * inputs, sizeof(float) * inputs);
        copy_out(c.forward(in), out + static_cast<size_t>(b) * outputs);
    }
}
}

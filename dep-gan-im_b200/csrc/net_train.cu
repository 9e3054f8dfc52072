// Training graphs (critic WGAN-GP, generator loss) -- placeholder until the backward kernels land.
#include "net.h"

struct Train {};

int train_alloc(depgan_net* h, Bump& b) { (void)h; (void)b; return 0; }

extern "C" {
int depgan_critic_grads(depgan_net*, depgan_net*, int, const float*, const float*, const float*, const float*, float*,
                        int, int, void*) {
  depgan_set_error("depgan_critic_grads: not implemented");
  return -3;
}
int depgan_gen_eval(depgan_net*, depgan_net*, depgan_net*, const float*, const float*, const float*, float, float*,
                    double*, int, int, void*) {
  depgan_set_error("depgan_gen_eval: not implemented");
  return -3;
}
int depgan_gen_grads(depgan_net*, depgan_net*, depgan_net*, const float*, const float*, const float*, float, float*,
                     double*, int, int, void*) {
  depgan_set_error("depgan_gen_grads: not implemented");
  return -3;
}
int depgan_gen_loss_finalize(float*, const double*, void*) {
  depgan_set_error("depgan_gen_loss_finalize: not implemented");
  return -3;
}
}

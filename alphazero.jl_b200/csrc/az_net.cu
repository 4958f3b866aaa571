// az_net.cu -- policy/value network forward (placeholder until the tcgen05 tower lands in this round).
#include "az_internal.h"

az_net* az_make_resnet(az_ctx* ctx, int game, const az_resnet_hp* hp, int* status) {
  (void)game; (void)hp;
  ctx->err = "ResNet forward not built yet";
  *status = AZ_EUNSUPPORTED;
  return nullptr;
}
az_net* az_make_simplenet(az_ctx* ctx, int game, const az_simplenet_hp* hp, int* status) {
  (void)game; (void)hp;
  ctx->err = "SimpleNet forward not built yet";
  *status = AZ_EUNSUPPORTED;
  return nullptr;
}

// Layout of the packed GeneralRenderingNetwork weights (O2345_RNET_PACK_FLOATS floats, every matrix stored [in][out]);
// written by o2345/rendering_network.py::GeneralRenderingNetwork.packed(), read by both view-blending kernels.
// Reference: reconstruction/models/rendering_network.py:40-73 (layer shapes).
#pragma once
#include "../../include/o2345.h"

namespace o2345 {
namespace rpack {

constexpr int CM = O2345_MAP_CH;   // 60 channels per pixel: rgb(3) + feat(56) + pad(1)
constexpr int NF = 59;

constexpr int P_D0W = 0;                    // ray_dir_fc[0]  [4][16]
constexpr int P_D0B = P_D0W + 64;           // [16]
constexpr int P_D1W = P_D0B + 16;           // ray_dir_fc[2]  [16][64]  (59 used)
constexpr int P_D1B = P_D1W + 1024;         // [64]
constexpr int P_B0W = P_D1B + 64;           // base_fc[0]     [193][64]: rows 0..15 geo, 16..74 mean, 75..133 var, 134..192 feat
constexpr int P_B0B = P_B0W + 193 * 64;     // [64]
constexpr int P_B1W = P_B0B + 64;           // base_fc[2]     [64][32]
constexpr int P_B1B = P_B1W + 2048;         // [32]
constexpr int P_V0W = P_B1B + 32;           // vis_fc[0]      [32][32]
constexpr int P_V0B = P_V0W + 1024;         // [32]
constexpr int P_V1W = P_V0B + 32;           // vis_fc[2]      [32][32]  residual outputs
constexpr int P_V1B = P_V1W + 1024;         // [32]
constexpr int P_V1V = P_V1B + 32;           // [32]      visibility output row
constexpr int P_V1VB = P_V1V + 32;          // [4]       its bias (first element)
constexpr int P_U0W = P_V1VB + 4;           // vis_fc2[0]     [32][32]
constexpr int P_U0B = P_U0W + 1024;         // [32]
constexpr int P_U1W = P_U0B + 32;           // vis_fc2[2]     [32]
constexpr int P_U1B = P_U1W + 32;           // [4]
constexpr int P_R0W = P_U1B + 4;            // rgb_fc[0]      [37][16]
constexpr int P_R0B = P_R0W + 592;          // [16]
constexpr int P_R1W = P_R0B + 16;           // rgb_fc[2]      [16][8]
constexpr int P_R1B = P_R1W + 128;          // [8]
constexpr int P_R2W = P_R1B + 8;            // rgb_fc[4]      [8]
constexpr int P_R2B = P_R2W + 8;            // [4]
constexpr int P_S = P_R2B + 4;              // [4]  |s| of the pooling weight
constexpr int P_TOTAL = P_S + 4;
static_assert(P_TOTAL == O2345_RNET_PACK_FLOATS, "header and kernels disagree on the rendering-net pack");

}  // namespace rpack
}  // namespace o2345

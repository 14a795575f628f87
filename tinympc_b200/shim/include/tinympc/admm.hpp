// forwarding header: the reference layout <tinympc/admm.hpp> -> the B200 shim
#pragma once
#include "../../tinympc_shim.hpp"

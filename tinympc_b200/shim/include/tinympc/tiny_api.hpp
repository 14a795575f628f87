// forwarding header: the reference layout <tinympc/tiny_api.hpp> -> the B200 shim
#pragma once
#include "../../tinympc_shim.hpp"

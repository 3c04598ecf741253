// oracle/compat/directx/d3d12.h — TEST INFRASTRUCTURE. Empty: the hot-path sources
// (DirectXTexP.h:142) include it but use nothing from it.
#pragma once

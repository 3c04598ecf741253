// oracle/compat/wsl/wrladapter.h — TEST INFRASTRUCTURE (DirectXTexP.h:141).
// Only the name Microsoft::WRL::ComPtr must exist ("using" declarations at
// DirectXTexConvert.cpp:17 etc.); it is never instantiated off-Windows.
#pragma once
namespace Microsoft { namespace WRL { template<typename T> class ComPtr; } }

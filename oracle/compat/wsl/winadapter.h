// oracle/compat/wsl/winadapter.h — TEST INFRASTRUCTURE (oracle build only).
// Minimal Windows-type surface that the hot-path sources of the reference need
// off-Windows (DirectXTex.h:34, DirectXTexP.h:140).  Constants are the public
// Win32 HRESULT ABI.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cmath>

typedef int32_t HRESULT;
typedef uint32_t UINT;
typedef uint32_t DWORD;
typedef int32_t LONG;
typedef int BOOL;
typedef uint8_t BYTE;
typedef uint16_t WORD;
typedef void* HANDLE;
typedef wchar_t WCHAR;
typedef const wchar_t* LPCWSTR;
typedef size_t SIZE_T;
typedef uint64_t UINT64;
typedef int64_t LONGLONG;

struct GUID { uint32_t Data1; uint16_t Data2; uint16_t Data3; uint8_t Data4[8]; };
typedef GUID IID;
#define REFGUID const GUID&
#define REFIID const IID&

#define S_OK            static_cast<HRESULT>(0)
#define S_FALSE         static_cast<HRESULT>(1)
#define E_NOTIMPL       static_cast<HRESULT>(0x80004001)
#define E_NOINTERFACE   static_cast<HRESULT>(0x80004002)
#define E_POINTER       static_cast<HRESULT>(0x80004003)
#define E_ABORT         static_cast<HRESULT>(0x80004004)
#define E_FAIL          static_cast<HRESULT>(0x80004005)
#define E_UNEXPECTED    static_cast<HRESULT>(0x8000FFFF)
#define E_ACCESSDENIED  static_cast<HRESULT>(0x80070005)
#define E_HANDLE        static_cast<HRESULT>(0x80070006)
#define E_OUTOFMEMORY   static_cast<HRESULT>(0x8007000E)
#define E_INVALIDARG    static_cast<HRESULT>(0x80070057)
#define E_BOUNDS        static_cast<HRESULT>(0x8000000B)

#define SUCCEEDED(hr) (static_cast<HRESULT>(hr) >= 0)
#define FAILED(hr)    (static_cast<HRESULT>(hr) < 0)

#define UNREFERENCED_PARAMETER(x) (void)(x)
#define __cdecl
#define __stdcall
#define WINAPI
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#ifndef _countof
#define _countof(a) (sizeof(a) / sizeof((a)[0]))
#endif
#ifndef UINT32_MAX
#define UINT32_MAX 0xffffffffu
#endif

// SAL annotations → nothing
#define _In_
#define _In_z_
#define _In_opt_
#define _In_opt_z_
#define _Out_
#define _Out_opt_
#define _Inout_
#define _Inout_opt_
#define _In_reads_(x)
#define _In_reads_opt_(x)
#define _In_reads_bytes_(x)
#define _In_reads_bytes_opt_(x)
#define _Out_writes_(x)
#define _Out_writes_opt_(x)
#define _Out_writes_all_(x)
#define _Out_writes_bytes_(x)
#define _Out_writes_bytes_opt_(x)
#define _Out_writes_bytes_to_(x, y)
#define _Out_writes_bytes_to_opt_(x, y)
#define _Out_writes_to_(x, y)
#define _Out_writes_to_opt_(x, y)
#define _Inout_updates_(x)
#define _Inout_updates_all_(x)
#define _Inout_updates_all_opt_(x)
#define _Inout_updates_bytes_(x)
#define _Inout_updates_bytes_all_(x)
#define _Inout_updates_bytes_all_opt_(x)
#define _In_range_(a, b)
#define _In_count_(x)
#define _In_bytecount_(x)
#define _Out_cap_(x)
#define _Out_bytecap_(x)
#define _Outptr_
#define _Outptr_opt_
#define _COM_Outptr_
#define _COM_Outptr_opt_
#define _Success_(x)
#define _When_(a, b)
#define _Reserved_
#define _Ret_maybenull_
#define _Ret_notnull_
#define _Use_decl_annotations_
#define _Analysis_assume_(x)
#define _Check_return_
#define _Null_terminated_
#define _Pre_null_
#define _Post_satisfies_(x)
#define _Field_size_(x)
#define _Field_size_opt_(x)
#define _Field_size_bytes_(x)

#include <cfloat>
#include <cmath>
using std::isnan;
using std::isinf;

// winnt.h DEFINE_ENUM_FLAG_OPERATORS (used by DirectXTex.inl:23-33)
#define DEFINE_ENUM_FLAG_OPERATORS(E) \
    extern "C++" { \
    inline constexpr E operator | (E a, E b) noexcept { return E(static_cast<uint32_t>(a) | static_cast<uint32_t>(b)); } \
    inline E& operator |= (E& a, E b) noexcept { return a = a | b; } \
    inline constexpr E operator & (E a, E b) noexcept { return E(static_cast<uint32_t>(a) & static_cast<uint32_t>(b)); } \
    inline E& operator &= (E& a, E b) noexcept { return a = a & b; } \
    inline constexpr E operator ~ (E a) noexcept { return E(~static_cast<uint32_t>(a)); } \
    inline constexpr E operator ^ (E a, E b) noexcept { return E(static_cast<uint32_t>(a) ^ static_cast<uint32_t>(b)); } \
    inline E& operator ^= (E& a, E b) noexcept { return a = a ^ b; } \
    }

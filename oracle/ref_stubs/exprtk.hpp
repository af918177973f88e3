// Shadows exprtk.hpp (a 40k-line third-party expression parser that mader_types.hpp only names in one struct).
#pragma once
namespace exprtk
{
template <typename T>
struct expression
{
};
}  // namespace exprtk

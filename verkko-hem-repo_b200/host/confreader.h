// Mirror of the reference's confreader (/root/reference/confreader/inc/confreader.h:109-118): declares the same
// 2 subsections / 37 entries with the same (mis)spellings and defaults (declare.cc:115-291) and parses a .prm file.
// GPU-only knobs are ADDITIVE entries with defaults, so an existing configuration.prm parses unchanged.
#ifndef VH_HOST_CONFREADER_H
#define VH_HOST_CONFREADER_H

#include "param_handler.h"

namespace vhhost
{
class confreader
{
public:
  explicit confreader(ParameterHandler &);
  void read_parameters(const std::string &);

private:
  ParameterHandler &prm;
  void              declare_parameters();
};
} // namespace vhhost
#endif

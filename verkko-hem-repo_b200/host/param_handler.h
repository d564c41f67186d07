// Minimal stand-in for dealii::ParameterHandler: the subset the reference's confreader and FemGL use
// (/root/reference/confreader/src/declare.cc:110-300, confreader.cc:111-121; femgl.cc:145-155, run.cc:116-133,
//  solve.cc:119-123, iteration.cc:119-122).  Same call names and meaning: declare_entry / enter_subsection /
// leave_subsection / get_double / get_integer / get_bool / parse_input.  The .prm syntax accepted is deal.II's
// ("subsection NAME" ... "end", "set KEY = VALUE", '#' comments); undeclared keys are an error, as in deal.II.
#ifndef VH_HOST_PARAM_HANDLER_H
#define VH_HOST_PARAM_HANDLER_H

#include <map>
#include <string>
#include <vector>

namespace vhhost
{
class ParameterHandler
{
public:
  void        declare_entry(const std::string &key, const std::string &default_value, const std::string &doc = "");
  void        enter_subsection(const std::string &name);
  void        leave_subsection();
  std::string get(const std::string &key) const;
  double      get_double(const std::string &key) const;
  long        get_integer(const std::string &key) const;
  bool        get_bool(const std::string &key) const;
  void        set(const std::string &key, const std::string &value);
  void        parse_input(const std::string &filename);
  void        parse_input_from_string(const std::string &text);
  // all "subsection/key" names in declaration order (tests compare them with the reference's key list)
  std::vector<std::string> declared_keys() const { return order; }

private:
  std::string                        path() const;
  std::vector<std::string>           stack;
  std::map<std::string, std::string> values;
  std::vector<std::string>           order;
};
} // namespace vhhost
#endif

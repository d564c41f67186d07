#include "param_handler.h"

#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace vhhost
{
namespace
{
std::string trim(const std::string &s)
{
  const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
} // namespace

std::string ParameterHandler::path() const
{
  std::string p;
  for (const auto &s : stack)
    p += s + "/";
  return p;
}
void ParameterHandler::declare_entry(const std::string &key, const std::string &def, const std::string &)
{
  const std::string k = path() + key;
  if (!values.count(k))
    order.push_back(k);
  values[k] = def;
}
void ParameterHandler::enter_subsection(const std::string &name) { stack.push_back(name); }
void ParameterHandler::leave_subsection()
{
  if (stack.empty())
    throw std::runtime_error("ParameterHandler::leave_subsection without enter_subsection");
  stack.pop_back();
}
std::string ParameterHandler::get(const std::string &key) const
{
  auto it = values.find(path() + key);
  if (it == values.end())
    throw std::runtime_error("ParameterHandler: entry <" + path() + key + "> was not declared");
  return it->second;
}
double ParameterHandler::get_double(const std::string &key) const
{
  const std::string v = get(key);
  char             *end = nullptr;
  const double      d = std::strtod(v.c_str(), &end);
  if (end == v.c_str() || *trim(end).c_str() != '\0')
    throw std::runtime_error("ParameterHandler: <" + key + "> = '" + v + "' is not a double");
  return d;
}
long ParameterHandler::get_integer(const std::string &key) const
{
  const std::string v = get(key);
  char             *end = nullptr;
  const long        d = std::strtol(v.c_str(), &end, 10);
  if (end == v.c_str() || *trim(end).c_str() != '\0')
    throw std::runtime_error("ParameterHandler: <" + key + "> = '" + v + "' is not an integer");
  return d;
}
bool ParameterHandler::get_bool(const std::string &key) const
{
  const std::string v = get(key);
  if (v == "true" || v == "yes" || v == "on")
    return true;
  if (v == "false" || v == "no" || v == "off")
    return false;
  throw std::runtime_error("ParameterHandler: <" + key + "> = '" + v + "' is not a bool");
}
void ParameterHandler::set(const std::string &key, const std::string &value)
{
  auto it = values.find(path() + key);
  if (it == values.end())
    throw std::runtime_error("ParameterHandler: cannot set undeclared entry <" + path() + key + ">");
  it->second = value;
}
void ParameterHandler::parse_input(const std::string &filename)
{
  std::ifstream in(filename);
  if (!in)
    throw std::runtime_error("ParameterHandler: cannot open " + filename);
  std::stringstream ss;
  ss << in.rdbuf();
  parse_input_from_string(ss.str());
}
void ParameterHandler::parse_input_from_string(const std::string &text)
{
  std::istringstream in(text);
  std::string        line;
  const size_t       depth0 = stack.size();
  int                lineno = 0;
  while (std::getline(in, line))
    {
      ++lineno;
      const size_t hash = line.find('#');
      if (hash != std::string::npos)
        line = line.substr(0, hash);
      line = trim(line);
      if (line.empty())
        continue;
      if (line.compare(0, 10, "subsection") == 0 && (line.size() == 10 || line[10] == ' ' || line[10] == '\t'))
        enter_subsection(trim(line.substr(10)));
      else if (line == "end")
        {
          if (stack.size() <= depth0)
            throw std::runtime_error("ParameterHandler: unbalanced 'end' at line " + std::to_string(lineno));
          leave_subsection();
        }
      else if (line.compare(0, 3, "set") == 0 && (line[3] == ' ' || line[3] == '\t'))
        {
          const size_t eq = line.find('=');
          if (eq == std::string::npos)
            throw std::runtime_error("ParameterHandler: missing '=' at line " + std::to_string(lineno));
          set(trim(line.substr(3, eq - 3)), trim(line.substr(eq + 1)));
        }
      else
        throw std::runtime_error("ParameterHandler: cannot parse line " + std::to_string(lineno) + ": " + line);
    }
  if (stack.size() != depth0)
    throw std::runtime_error("ParameterHandler: unbalanced subsection/end");
}
} // namespace vhhost

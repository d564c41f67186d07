// VerHem-compatible driver (/root/reference/sol/src/main.cc:101-148): reads "configuration.prm" from the working
// directory (or argv[1]), constructs FemGL<3>, runs it; any exception prints a banner and returns 1.
#include <iostream>

#include "confreader.h"
#include "femgl.h"

int main(int argc, char *argv[])
{
  try
    {
      using namespace vhhost;
      ParameterHandler prm;
      confreader       cr(prm);
      cr.read_parameters(argc > 1 ? argv[1] : "configuration.prm");
      prm.enter_subsection("control parameters");
      const unsigned int degree = (unsigned int)prm.get_integer("polynomial degree");
      prm.leave_subsection();
      FemGL<3> femgl(degree, prm);
      femgl.run();
    }
  catch (std::exception &exc)
    {
      std::cerr << std::endl
                << std::endl
                << "----------------------------------------------------" << std::endl;
      std::cerr << "Exception on processing: " << std::endl
                << exc.what() << std::endl
                << "Aborting!" << std::endl
                << "----------------------------------------------------" << std::endl;
      return 1;
    }
  catch (...)
    {
      std::cerr << std::endl
                << std::endl
                << "----------------------------------------------------" << std::endl;
      std::cerr << "Unknown exception!" << std::endl
                << "Aborting!" << std::endl
                << "----------------------------------------------------" << std::endl;
      return 1;
    }
  return 0;
}

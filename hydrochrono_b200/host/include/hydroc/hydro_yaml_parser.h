// hydro.yaml reader (reference src/hydro_yaml_parser.h:20).
#ifndef HYDROC_B200_HYDRO_YAML_PARSER_H
#define HYDROC_B200_HYDRO_YAML_PARSER_H

#include <string>

#include <hydroc/hydro_types.h>

// Throws std::runtime_error("Could not open hydro file: ...") / ("No 'hydrodynamics:' section found ...") and the
// waves.period / waves validation errors of the reference.
YAMLHydroData ReadHydroYAML(const std::string& hydro_file_path);

#endif

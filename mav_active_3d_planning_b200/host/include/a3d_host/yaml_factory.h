// yaml_factory.h — ModuleFactory over a YAML file, walking it the way ModuleFactoryROS walks the ROS
// parameter server (active_3d_planning_ros/src/module/module_factory_ros.cpp:17-56): the direct scalar
// children of a namespace become strings, "type" selects the class, "param_namespace" records the
// namespace, booleans are stringified as 1/0 (XmlRpc's operator<<), nested maps are child namespaces.
// The reference's planner YAMLs (app_reconstruction/cfg/planners/*.yaml) load unchanged.
#pragma once

#include <map>
#include <string>
#include <vector>

#include "a3d_host/module.h"

namespace active_3d_planning {

// The YAML subset the planner configs use: nested block mappings, scalar values, comments.
class YamlTree {
 public:
  // key → scalar (flattened with '/' separators, absolute paths "/a/b/c")
  bool loadFile(const std::string& path, std::string* error);
  bool loadString(const std::string& text, std::string* error);
  // mount the tree below a prefix namespace (e.g. "/planner_node")
  void setRoot(const std::string& ns) { root_ = ns; }
  bool hasNamespace(const std::string& ns) const;
  // direct scalar children of ns
  std::map<std::string, std::string> scalars(const std::string& ns) const;
  void set(const std::string& path, const std::string& value) { values_[normalize(path)] = value; }
  const std::map<std::string, std::string>& all() const { return values_; }

 private:
  std::string normalize(const std::string& ns) const;
  std::map<std::string, std::string> values_;
  std::string root_;
};

class ModuleFactoryYaml : public ModuleFactory {
 public:
  explicit ModuleFactoryYaml(const YamlTree& tree) : tree_(tree) {}
  bool getParamMapAndType(Module::ParamMap* map, std::string* type, std::string args) override;
  std::vector<std::string> verboseLog;  // printVerbose output, newest last

 protected:
  void printVerbose(const Module::ParamMap& map) override;
  void printError(const std::string& message) override;
  YamlTree tree_;
};

}  // namespace active_3d_planning

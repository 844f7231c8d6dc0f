// module.h — module system of the view-gain path: string-typed parameter maps, self-registering
// module types, and the factory that builds a module from a namespace of a config store.
//
// Mirrors the reference's interface so that modules written against it port unchanged:
//   ModuleBase / Module                active_3d_planning_core/include/active_3d_planning_core/module/module.h:23-86
//   ModuleFactoryRegistry              .../module/module_factory_registry.h:24-80
//   ModuleFactory                      .../module/module_factory.h:17-68, src/module/module_factory.cpp:14-41
//   PlannerI                           .../planner/planner_I.h:22-55
// Differences: unknown / duplicate type names throw std::runtime_error via a3d_fatal() where the
// reference calls glog's LOG(FATAL) (module_factory_registry.h:47-50, :65-69).
#pragma once

#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeindex>
#include <typeinfo>

#include "a3d_host/eigen_lite.h"

namespace active_3d_planning {

class Module;
class ModuleFactory;
class Map;
class TrajectoryEvaluator;
struct SystemConstraints;

[[noreturn]] void a3d_fatal(const std::string& message);

// What a module may ask of its owner (the subset of planner_I.h this path touches; the generator /
// back-tracker accessors of the reference belong to out-of-scope subsystems).
class PlannerI {
 public:
  virtual ~PlannerI() = default;
  virtual const Eigen::Vector3d& getCurrentPosition() const = 0;
  virtual const Eigen::Quaterniond& getCurrentOrientation() const = 0;
  virtual TrajectoryEvaluator& getTrajectoryEvaluator() = 0;
  virtual ModuleFactory& getFactory() = 0;
  virtual Map& getMap() = 0;
  virtual SystemConstraints& getSystemConstraints() = 0;
  virtual void printInfo(const std::string& text) = 0;
  virtual void printWarning(const std::string& text) = 0;
  virtual void printError(const std::string& text) = 0;
};

class ModuleBase {
 public:
  typedef std::map<std::string, std::string> ParamMap;
  ModuleBase() : verbose_modules_(false) {}
  virtual ~ModuleBase() = default;
  virtual void setupFromParamMap(ParamMap* param_map) = 0;
  virtual void setVerbose(bool verbose) { verbose_modules_ = verbose; }
  // throws std::invalid_argument with the module's message if checkParamsValid fails
  void assureParamsValid();
  virtual bool checkParamsValid(std::string* /*error_message*/) { return true; }

 protected:
  // string → T with `istringstream >> T` (bools therefore read "1"/"0"); appends one line per
  // parameter to map["verbose_text"], "<name>: <value>" or "<name>: <default> (default)"
  template <typename T>
  void setParam(ParamMap* map, std::string param_name, T* param, T param_default) {
    std::string text;
    ParamMap::iterator it = map->find("verbose_text");
    if (it != map->end()) text = it->second + "\n";
    text += param_name + ": ";
    it = map->find(param_name);
    if (it != map->end()) {
      std::istringstream iss(it->second);
      iss >> *param;
      text += it->second;
    } else {
      *param = param_default;
      std::ostringstream oss;
      oss << param_default;
      text += oss.str() + " (default)";
    }
    (*map)["verbose_text"] = text;
  }
  bool verbose_modules_;
};

class Module : public ModuleBase {
 public:
  explicit Module(PlannerI& planner) : planner_(planner) {}  // NOLINT
  virtual ~Module() = default;

 protected:
  PlannerI& planner_;
};

using ModuleFactoryMethod = std::function<Module*(PlannerI& planner)>;

// name → constructor table filled by static Registration objects
class ModuleFactoryRegistry {
 public:
  static ModuleFactoryRegistry& Instance();

  template <class ModuleType>
  std::unique_ptr<ModuleType> createModule(const std::string& module_name, PlannerI& planner,
                                           bool /*verbose*/) {
    auto it = getRegistryMap().find(module_name);
    if (it == getRegistryMap().end()) {
      a3d_fatal("Could not find type '" + module_name +
                "' to create a module from. Available types are: " + listTypes() + ".");
    }
    return std::unique_ptr<ModuleType>(dynamic_cast<ModuleType*>(it->second(planner)));
  }

  template <class ModuleType>
  struct Registration {
    explicit Registration(const std::string& module_name) {
      if (!getRegistryMap()
               .emplace(module_name, [](PlannerI& planner) -> Module* { return new ModuleType(planner); })
               .second) {
        a3d_fatal("ModuleFactory: Cannot register already existing module name '" + module_name + "'!");
      }
    }
  };

  static std::string listTypes();
  static bool isRegistered(const std::string& module_name);

 private:
  typedef std::map<std::string, ModuleFactoryMethod> ModuleMap;
  static ModuleMap& getRegistryMap();
};

// Builds modules from "args" (a namespace path of the config store).  Concrete factories say how a
// namespace becomes a ParamMap (ROS parameter server in the reference; a YAML file here).
class ModuleFactory {
 public:
  virtual ~ModuleFactory() = default;

  template <class ModuleType>
  std::unique_ptr<ModuleType> createModule(const std::string& args, PlannerI& planner, bool verbose) {
    Module::ParamMap param_map;
    std::string type = "NoDefaultTypeAvailable";
    getDefaultType(std::type_index(typeid(ModuleType)), &type);
    getParamMapAndType(&param_map, &type, args);
    auto module = ModuleFactoryRegistry::Instance().createModule<ModuleType>(type, planner, verbose);
    if (!module) a3d_fatal("module type '" + type + "' is not of the requested kind (" + args + ")");
    module->setVerbose(verbose);
    module->setupFromParamMap(&param_map);
    module->assureParamsValid();
    if (verbose) printVerbose(param_map);
    return module;
  }

  void registerLinkableModule(const std::string& name, Module* module);
  Module* readLinkableModule(const std::string& name);

  virtual bool getParamMapAndType(Module::ParamMap* map, std::string* type, std::string args) = 0;

 protected:
  // types used when a namespace names none (module_factory.cpp:14-25): BoundingVolume only here
  void getDefaultType(const std::type_index& module_index, std::string* type);
  virtual void printVerbose(const Module::ParamMap& map) = 0;
  virtual void printError(const std::string& message) = 0;
  std::map<std::string, Module*> linkable_module_list_;
};

}  // namespace active_3d_planning

classdef tenscalc_b200_solver < handle
% Batched counterpart of the class cmex2optimizeCS generates (lib/cmex2optimizeCS.m:280-294;
% doc/sphinx/optimization.rst:487-556).  Same method names; every value carries the batch as
% trailing dimension.  `image` is the problem image written by tenscalc_b200.tape.serialize
% (or, once the .cg reader of SURVEY.md 8(f) lands, by saveScalarized).
%
%   obj=tenscalc_b200_solver('tmp_quadcopter.tcimg',65536,0);
%   setP(obj,'Ts',.02); setV(obj,'p',p0);        % also callable as setP_Ts(obj,.02) via subsref
%   [status,iter,time]=solve(obj,1e-3,int32(1000),int32(-1));
%   y=getOutputs(obj);
    properties (SetAccess=private)
        h; names; batch;
    end
    methods
        function obj=tenscalc_b200_solver(imageFile,batch,device,moduleFile)
            fid=fopen(imageFile,'r'); img=fread(fid,inf,'*uint8'); fclose(fid);
            if nargin<4   % tape interpreter; pass the compiled module (tenscalc_b200/_gen/*.so) for native code
                obj.h=tenscalc_b200_mex('create',img,batch,device);
            else
                obj.h=tenscalc_b200_mex('create',img,batch,device,moduleFile);
            end
            obj.batch=batch;
            obj.names=jsondecode(fileread([imageFile,'.json'])); % {sets:{..},outputs:{..}}
        end
        function setP(obj,name,value); obj.set(['P_',name],value); end
        function setV(obj,name,value); obj.set(['V_',name],value); end
        function set(obj,key,value)
            % by name: the image carries the <name> of the reference's <class>_set_<name> symbols (tc_find_set)
            k=key(3:end);
            tenscalc_b200_mex('set',obj.h,k,reshape(double(value),[],size(value,ndims(value))*(numel(value)>1)+(numel(value)==1)));
        end
        function [status,iter,time]=solve(obj,mu0,maxIter,saveIter,addEye2Hessian)
            if nargin<5; addEye2Hessian=[1e-9;1e-9]; end
            if nargin<4; saveIter=int32(-1); end
            [status,iter,time]=tenscalc_b200_mex('solve',obj.h,mu0,maxIter,saveIter,addEye2Hessian);
        end
        function varargout=getOutputs(obj)
            y=tenscalc_b200_mex('getOutputs',obj.h);
            if nargout<=1; varargout{1}=cell2struct(y(:),obj.names.outputs(:),1); else; varargout=y; end
        end
        function varargout=subsref(obj,s)   % setP_<name>(obj,v) / setV_<name>(obj,v) spelling
            if s(1).type=='.' && numel(s(1).subs)>5 && any(strcmp(s(1).subs(1:5),{'setP_','setV_'}))
                obj.set(s(1).subs(4:end),s(2).subs{:});
            else
                [varargout{1:nargout}]=builtin('subsref',obj,s);
            end
        end
        function delete(obj); tenscalc_b200_mex('delete',obj.h); end
    end
end
